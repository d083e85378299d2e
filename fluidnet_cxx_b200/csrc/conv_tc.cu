// conv_tc.cu -- MultiScaleNet forward on the sm_100a tensor cores (tcgen05 + TMEM), im2col-free.
//
// Reference: pytorch/lib/multi_scale_net.py:101-127 (the 17 nn.Conv2d of the three-scale pyramid).
// Every 3x3 and 5x5 layer runs here as an implicit GEMM
//     D[pixels x Cout] += A[pixels x Cin] . W_tap[Cin x Cout]      per filter tap
// (channel counts are zero-padded to MMA granularity: Cin to a multiple of 16, Cout to 16/32/64/128).
//
// fp32 parity through fp16 tensor cores (the parity bar is 1e-5 relative):
// every activation a and weight w is carried as an exact two-term fp16 expansion of a power-of-two
// scaled value,  s*a = hi + lo  (22 significant bits), and the product is formed as
// hi*hi + hi*lo + lo*hi in three kind::f16 MMAs with fp32 TMEM accumulation (the dropped lo*lo term
// is 2^-22 relative).  The tensor core's fp32 accumulate truncates (round toward zero), a bias of
// half an ulp per accumulation: the two small cross terms therefore go to a SEPARATE "corr"
// accumulator (their truncation ulps are 2^-11 smaller) and only the hi*hi term touches the main
// one; the epilogue adds the two in round-to-nearest fp32.  Scales are powers of two, so scaling and
// un-scaling are exact.  The activation scale of a layer's OUTPUT is chosen before the layer runs
// from the rigorous bound  max|y| <= max|x| * max_n sum_k |w_nk| + max|b|  (max|x| is measured by the
// producing kernel with an atomicMax), so fp16 can never overflow and the resolution floor stays
// below 2^-30 of the bound.
//
// Data layout in HBM ("split chunked"): two fp16 planes (hi, lo), each [C/8][H+2P][W+2P][8 halves]
// with a zero border of P = 2 pixels that no kernel ever writes (it IS the convolution's zero
// padding).  One (pixel, 8-channel chunk) = 16 bytes = one row of a K-major UMMA core matrix, so a
// staged row of pixels is directly a no-swizzle K-major operand whose row index is affine in the
// pixel index (SBO = 128 B): the A operand of filter tap (ky, kx) is the SAME shared-memory tile with
// the descriptor start address advanced by (ky*row_pitch + kx)*16 bytes (validated on hardware by
// tools/tc_probe.cu).  Each input row is staged once per 16-channel chunk by 1-D bulk copies of the
// TMA engine (cp.async.bulk, SASS UBLKCP) and reused by all KS*KS taps and all 3 split terms.
//
// Kernel shape: persistent, one CTA per SM, 320 threads = producer warp (bulk copies), MMA warp (one
// lane issues tcgen05.mma), 8 epilogue warps (tcgen05.ld -> bias/ReLU -> split -> 16-byte stores).
// A block of work = R output rows x 128 pixels x Cout: per M tile a [main | corr] accumulator pair of
// 128 x 2*Cout fp32 in TMEM (2*R*Cout <= 512 columns), so each streamed weight slot is reused by R M
// tiles.  Per tap two MMAs: a_hi.[w_hi;w_lo] (N = 2*Cout, fills main and corr at once) and a_lo.w_hi.
// The first and last 16-channel chunk of a block run tile-major with per-tile TMEM barriers so the
// epilogue of one tile overlaps the MMAs of the next.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "../../include/fluidstep.h"
#include "cnn_internal.h"
#include "host_util.h"
#include "tc_ptx.cuh"

namespace fnx {
namespace tc {

constexpr int PAD = FNX_TC_PAD;  // zero border of the split chunked layout
constexpr int TW = 128;          // pixels per M tile (one row segment)

struct ActMeta {  // == fnx_act_meta
  unsigned amax_bits;  // fp32 bit pattern of max|a| over the tensor (atomicMax target)
  float scale;         // power of two the stored hi+lo expansion is multiplied by
};

// power of two s with bound*s <= 2^14 (fp16 max is 2^16 - 32)
__host__ __device__ __forceinline__ float pow2_scale_for(float bound) {
  if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
  int e;
  frexpf(bound, &e);  // bound = f * 2^e, f in [0.5, 1)
  int k = 14 - e;
  k = k > 100 ? 100 : (k < -100 ? -100 : k);
  return ldexpf(1.f, k);
}

struct ConvArgs {
  const __half* x;     // input planes: hi at x, lo at x + x_plane
  size_t x_plane;      // halves
  const uint8_t* w;    // packed weights: [Cin_pad/16][KS (ky)] slots, slot = [kx][j][plane][COUT][8 halves]
  const float* bias;   // Cout real values
  void* y;             // split planes (out_mode 0) or fp32 NCHW (out_mode 1)
  size_t y_plane;      // halves (out_mode 0)
  const ActMeta* in_meta;
  ActMeta* out_meta;
  int nchunks;         // Cin_pad / 16
  int Cout;            // real output channels (<= COUT)
  int H, W, Hp, Wp;
  int relu, out_mode, y_ctotal, y_coff;
  float w_scale, w_norm, b_max;
  int tiles_x, nblocks;
  int R;               // output rows per block (<= the kernel's RMAX template parameter)
  int iters;           // blocks per CTA in cluster mode (same for every CTA)
  long long* dbg;      // optional (fnx_tc_set_debug): per CTA {A_FULL wait, W_FULL wait, ACC_EMPTY wait, total} clocks
                       // of the MMA warp
  // out_mode 2: fused 1x1 head (multi_scale_net.py:116 `final`): y[pixel] = sum_c head_w[c]*out[c] + head_b,
  // one fp32 channel, for layers with Cout <= 16 (the 32->8 5x5 layer feeding the 8->1 conv)
  const float* head_w;
  const float* head_b;
  // the packed weights may be replicated `w_nrep` times, `w_rep_stride` bytes apart: CTA b streams copy
  // b % w_nrep, which spreads the (identical, simultaneous) weight fills of all SMs over more L2 lines
  int w_nrep;
  size_t w_rep_stride;
  // A BATCH of images as one tall image: image n occupies rows [n*slice_pitch, n*slice_pitch + slice_h) of the H rows
  // (slice_pitch = slice_h + 2*PAD: between two images lie the zero pad rows of both, which no kernel writes, so every
  // image keeps its own zero padding).  One launch per layer for the whole batch; rows in the gaps are computed and
  // dropped.  A single image is the case slice_h = H.
  int slice_h, slice_pitch;
};

constexpr int NEPI = 8;                      // epilogue warps (two per TMEM lane quadrant)
constexpr int NTHREADS = 32 * (2 + NEPI);

template <int KS, int COUT, int R, int WS>
struct Cfg {
  static constexpr int HALO = KS / 2;
  static constexpr int RP = TW + KS - 1;              // staged row pitch in pixels
  static constexpr int ROWS = R + KS - 1;             // staged rows per block
  static constexpr uint32_t ROWB = RP * 16;           // bytes of one staged row of one 8-channel chunk
  static constexpr uint32_t A_J = ROWS * ROWB;        // stride between the two 8-channel chunks = LBO(A)
  static constexpr uint32_t A_PLANE = 2 * A_J;
  static constexpr uint32_t A_STAGE = 2 * A_PLANE;    // hi + lo
  static constexpr uint32_t W_J = 2 * COUT * 16;      // [w_hi rows ; w_lo rows] of one 8-channel chunk = LBO(B)
  static constexpr uint32_t W_KX = 2 * W_J;
  static constexpr uint32_t W_STAGE = KS * W_KX;      // one (16-channel chunk, ky) slot
  // activation stages: a stage can only be refilled once every MMA that read it has completed, so a
  // third stage lets the fill of chunk c+2 start while chunk c is still running (when it fits)
  static constexpr uint32_t FIXED = WS * W_STAGE + (6 + 2 * WS + 2 * R) * 8 + 16 + COUT * 4;
  static constexpr int AS = (3 * A_STAGE + FIXED <= 227 * 1024) ? 3 : 2;
  static constexpr uint32_t NBAR = 2 * AS + 2 * WS + 2 * R;
  static constexpr uint32_t SMEM = AS * A_STAGE + WS * W_STAGE + NBAR * 8 + 16 + COUT * 4;
  static constexpr uint32_t NCOLS = 2 * R * COUT;     // per M tile: [main | corr] accumulators
  static constexpr uint32_t TMEM_COLS = NCOLS <= 32 ? 32 : NCOLS <= 64 ? 64 : NCOLS <= 128 ? 128 : NCOLS <= 256 ? 256 : 512;
  static_assert(NCOLS <= 512, "accumulators exceed TMEM");
  static_assert(WS >= KS + 1, "the first / last chunk of a block needs all KS weight slots of the chunk resident");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// One filter tap of one M tile: two MMAs.
//   [main | corr] += a_hi . [w_hi ; w_lo]^T      (N = 2*COUT: hi*hi into main, hi*lo into corr)
//          corr   += a_lo . w_hi^T               (N = COUT)
// The cross terms never touch the main accumulator (see the header on truncation bias).
template <int COUT>
__device__ __forceinline__ void mma_tap(uint32_t d_tile, uint64_t a_hi, uint64_t a_lo, uint64_t w, uint32_t accumulate) {
  mma_f16_ss(d_tile, a_hi, w, idesc_f16_f32acc(128, 2 * COUT), accumulate);
  mma_f16_ss(d_tile + COUT, a_lo, w, idesc_f16_f32acc(128, COUT), 1);
}

// CL = CTAs per cluster.  CL = 2 (opt-in, FNX_TC_CLUSTER=1): the two CTAs of a cluster work on different
// blocks but consume the same weight slots in lockstep, so rank 0 loads every slot ONCE with a multicast
// bulk copy into both CTAs' rings; a slot is refilled when both MMA issuers have committed it (multicast
// arrive on both W_EMPTY barriers, count 2).  Halves the L2->SM weight fills (DESIGN.md "Fill budget").
template <int KS, int COUT, int R, int WS, int CL>
__global__ void __launch_bounds__(NTHREADS, 1) k_conv_tc(const ConvArgs a) {
  using C = Cfg<KS, COUT, R, WS>;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  // every CTA of a cluster runs the same number of blocks (a CTA past the end re-runs the last block
  // without storing anything) so the weight-slot sequence stays in lockstep
  const int my_iters = CL > 1 ? a.iters : (a.nblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  constexpr int AS = C::AS;
  uint8_t* sW = smem + AS * C::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WS * C::W_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto A_FULL = [&](int s) { return bar0 + 8u * s; };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (AS + s); };
  auto W_FULL = [&](int s) { return bar0 + 8u * (2 * AS + s); };
  auto W_EMPTY = [&](int s) { return bar0 + 8u * (2 * AS + WS + s); };
  auto ACC_FULL = [&](int r) { return bar0 + 8u * (2 * AS + 2 * WS + r); };
  auto ACC_EMPTY = [&](int r) { return bar0 + 8u * (2 * AS + 2 * WS + R + r); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = a.nchunks;
  const uint8_t* const wsrc = a.w + (a.w_nrep > 1 ? (size_t)(blockIdx.x % a.w_nrep) * a.w_rep_stride : 0);
  const int Rrt = a.R;  // rows per block at run time (R = capacity the shared/tensor memory is sized for)
  const bool resident = nchunks * KS <= WS;  // all weight slots of the layer fit in the ring

  if (threadIdx.x == 0) {
    for (int s = 0; s < AS; s++) {
      mbar_init(A_FULL(s), 1);
      mbar_init(A_EMPTY(s), 1);
    }
    for (int s = 0; s < WS; s++) {
      mbar_init(W_FULL(s), 1);
      mbar_init(W_EMPTY(s), CL);
    }
    for (int r = 0; r < R; r++) {
      mbar_init(ACC_FULL(r), 1);
      mbar_init(ACC_EMPTY(r), NEPI);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = (a.bias && i < a.Cout) ? a.bias[i] : 0.f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();  // the peer's barriers exist before any multicast copy / arrive targets them
  const uint32_t taddr = *tmem_slot;
  // Programmatic dependent launch: everything above (and the weight prefetch below) touches only
  // static data and this CTA's own shared / tensor memory, so it may overlap the tail of the previous
  // kernel in the stream; pdl_wait() orders every read of upstream activations / metas after it.
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== producer: stage activation rows and weight slots with the TMA engine =====
    // Lane 0 owns the barrier protocol (waits, expect_tx); the bulk copies of a stage -- one per
    // (plane, 8-channel chunk, row), 2 KB each -- are issued by all 32 lanes in parallel: issued from a
    // single thread they cost more than the copies themselves on the narrow layers.
    int as = 0, aph = 0, ws = 0, wph = 0;
    auto load_w = [&](int slot_smem, int slot_gmem) {  // arm my barrier; one copy per CTA, or one per cluster
      mbar_expect_tx(W_FULL(slot_smem), C::W_STAGE);
      const uint8_t* src = wsrc + (size_t)slot_gmem * C::W_STAGE;
      if (CL == 1) bulk_g2s(smem_u32(sW + slot_smem * C::W_STAGE), src, C::W_STAGE, W_FULL(slot_smem));
      else if (crank == 0)
        bulk_g2s_multicast(smem_u32(sW + slot_smem * C::W_STAGE), src, C::W_STAGE, W_FULL(slot_smem), (uint16_t)((1u << CL) - 1));
    };
    if (resident && lane == 0) {  // the whole layer's weights fit in the ring: load once, never release
      for (int s = 0; s < nchunks * KS; s++) load_w(s, s);
    }
    pdl_wait();
    for (int it = 0; it < my_iters; it++) {
      int blk = (int)blockIdx.x + it * (int)gridDim.x;
      blk = blk < a.nblocks ? blk : a.nblocks - 1;
      const int x0 = (blk % a.tiles_x) * TW, y0 = (blk / a.tiles_x) * Rrt;
      // staged rows = padded-image rows y0+PAD-HALO ..; rows past the padded image are skipped
      int nrows = a.Hp - (y0 + PAD - C::HALO);
      nrows = nrows > Rrt + KS - 1 ? Rrt + KS - 1 : nrows;
      for (int c = 0; c < nchunks; c++) {
        if (lane == 0) {
          mbar_wait(A_EMPTY(as), aph ^ 1);
          mbar_expect_tx(A_FULL(as), (uint32_t)nrows * 4u * C::ROWB);
        }
        __syncwarp();
        const uint32_t dst0 = smem_u32(sA + as * C::A_STAGE);
        for (int e = lane; e < 4 * nrows; e += 32) {
          const int pj = e / nrows, r = e - pj * nrows;  // pj = plane * 2 + j
          const int pl = pj >> 1, j = pj & 1;
          const __half* src = a.x + pl * a.x_plane +
                              (((size_t)(c * 2 + j) * a.Hp + (y0 + PAD - C::HALO + r)) * a.Wp + (x0 + PAD - C::HALO)) * 8;
          bulk_g2s(dst0 + pl * C::A_PLANE + j * C::A_J + r * C::ROWB, src, C::ROWB, A_FULL(as));
        }
        if (!resident && lane == 0) {
          for (int ky = 0; ky < KS; ky++) {
            mbar_wait(W_EMPTY(ws), wph ^ 1);  // every CTA of the cluster has consumed this slot
            load_w(ws, c * KS + ky);
            if (++ws == WS) { ws = 0; wph ^= 1; }
          }
        }
        __syncwarp();
        if (++as == AS) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues =====
    // Chunk order: tile-major when the chunk is the last of its block (tile r's accumulators are
    // committed to the epilogue one tile at a time, so the epilogue of one tile overlaps the MMAs of
    // the next) or when the weights are resident; otherwise ky-major, which releases weight slots
    // progressively.  The first chunk re-acquires tile r from the epilogue just before touching it.
    int as = 0, aph = 0, ws = 0, wph = 0;
    long long t_a = 0, t_w = 0, t_acc = 0;
    const long long t_begin = a.dbg ? clock64() : 0;
    auto timed_wait = [&](uint32_t bar, uint32_t parity, long long& acc) {
      if (a.dbg) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
      } else {
        mbar_wait(bar, parity);
      }
    };
    auto release_w = [&](int slot) {
      if (CL == 1) mma_commit(W_EMPTY(slot));
      else mma_commit_multicast(W_EMPTY(slot), (uint16_t)((1u << CL) - 1));
    };
    for (int it = 0; it < my_iters; it++) {
      for (int c = 0; c < nchunks; c++) {
        const bool first = c == 0, last = c == nchunks - 1;
        timed_wait(A_FULL(as), aph, t_a);
        const uint32_t a_base = smem_u32(sA + as * C::A_STAGE);
        const uint64_t a_hi0 = smem_desc_kmajor_noswz(a_base, C::A_J, 128);
        const uint64_t a_lo0 = smem_desc_kmajor_noswz(a_base + C::A_PLANE, C::A_J, 128);
        if (resident || last) {
          uint64_t wd[KS];
          {
            int s = resident ? c * KS : ws, ph = resident ? 0 : wph;
#pragma unroll
            for (int ky = 0; ky < KS; ky++) {
              timed_wait(W_FULL(s), ph, t_w);
              wd[ky] = smem_desc_kmajor_noswz(smem_u32(sW + s * C::W_STAGE), C::W_J, 128);
              if (++s == WS) { s = 0; ph ^= 1; }
            }
          }
          tc_fence_after();
#pragma unroll
          for (int r = 0; r < R; r++) {
            if (r >= Rrt) break;
            if (first) {
              timed_wait(ACC_EMPTY(r), (it & 1) ^ 1, t_acc);
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int ky = 0; ky < KS; ky++)
#pragma unroll
                for (int kx = 0; kx < KS; kx++) {
                  const uint32_t aoff = (uint32_t)(((r + ky) * C::RP + kx) * 16) >> 4;
                  mma_tap<COUT>(taddr + (uint32_t)(r * 2 * COUT), a_hi0 + aoff, a_lo0 + aoff,
                                wd[ky] + ((uint32_t)(kx * C::W_KX) >> 4), (first && ky == 0 && kx == 0) ? 0u : 1u);
                }
              if (last) mma_commit(ACC_FULL(r));
            }
            __syncwarp();
          }
          if (!resident) {
#pragma unroll
            for (int ky = 0; ky < KS; ky++) {
              if (elect_one()) release_w(ws);
              if (++ws == WS) { ws = 0; wph ^= 1; }
            }
          }
        } else {
          for (int ky = 0; ky < KS; ky++) {
            timed_wait(W_FULL(ws), wph, t_w);
            tc_fence_after();
            const uint64_t w0 = smem_desc_kmajor_noswz(smem_u32(sW + ws * C::W_STAGE), C::W_J, 128);
#pragma unroll
            for (int r = 0; r < R; r++) {
              if (r >= Rrt) break;
              if (first && ky == 0) {
                timed_wait(ACC_EMPTY(r), (it & 1) ^ 1, t_acc);
                tc_fence_after();
              }
              if (elect_one()) {
#pragma unroll
                for (int kx = 0; kx < KS; kx++) {
                  const uint32_t aoff = (uint32_t)(((r + ky) * C::RP + kx) * 16) >> 4;
                  mma_tap<COUT>(taddr + (uint32_t)(r * 2 * COUT), a_hi0 + aoff, a_lo0 + aoff,
                                w0 + ((uint32_t)(kx * C::W_KX) >> 4), (first && ky == 0 && kx == 0) ? 0u : 1u);
                }
              }
              __syncwarp();
            }
            if (elect_one()) release_w(ws);
            __syncwarp();
            if (++ws == WS) { ws = 0; wph ^= 1; }
          }
        }
        if (elect_one()) mma_commit(A_EMPTY(as));
        __syncwarp();
        if (++as == AS) { as = 0; aph ^= 1; }
      }
    }
    if (a.dbg && lane == 0) {
      long long* d = a.dbg + 4 * (size_t)blockIdx.x;
      d[0] = t_a; d[1] = t_w; d[2] = t_acc; d[3] = clock64() - t_begin;
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias / ReLU -> (split fp16 | fp32 NCHW) =====
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half_id = (warp - 2) >> 2;  // which half of the 16-column groups this warp handles
    pdl_wait();
    const float s_in = a.in_meta->scale;
    const float inv = 1.f / (s_in * a.w_scale);
    const float s_out = pow2_scale_for(__uint_as_float(a.in_meta->amax_bits) * a.w_norm + a.b_max);
    float amax = 0.f;
    constexpr int NGRP = COUT / 16;
    for (int it = 0; it < my_iters; it++) {
      const int blk = (int)blockIdx.x + it * (int)gridDim.x;
      const bool live = blk < a.nblocks;  // a lockstep filler iteration computes but never stores
      const int x0 = (blk % a.tiles_x) * TW, y0 = (blk / a.tiles_x) * Rrt;
      const int px = x0 + q * 32 + lane;
#pragma unroll 1
      for (int r = 0; r < Rrt; r++) {
        const int y = y0 + r;
        const int img = y / a.slice_pitch, yy = y - img * a.slice_pitch;   // image of the batch, row inside it
        const bool ok = live && (y < a.H) && (px < a.W) && (yy < a.slice_h);
        mbar_wait(ACC_FULL(r), it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int g = half_id; g < NGRP; g += 2) {
          const int c0 = g * 16;
          if (c0 >= a.Cout) break;  // padded output channels (warp-uniform)
          uint32_t vm[16], vc[16];
          const uint32_t t0 = taddr + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * 2 * COUT + c0);
          tmem_ld16(t0, vm);
          tmem_ld16(t0 + (uint32_t)COUT, vc);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; i++) {
            float t = fmaf(__uint_as_float(vm[i]) + __uint_as_float(vc[i]), inv, s_bias[c0 + i]);
            if (a.relu) t = fmaxf(t, 0.f);
            f[i] = t;
          }
          if (ok) {  // lanes outside the image hold garbage accumulators: never stored, never in amax
            if (a.out_mode == 0) {
#pragma unroll
              for (int i = 0; i < 16; i++) amax = fmaxf(amax, fabsf(f[i]));
              __half* yh = reinterpret_cast<__half*>(a.y);
#pragma unroll
              for (int gg = 0; gg < 2; gg++) {
                __align__(16) __half2 hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  // two-term fp16 expansion of a pair with the packed conversions (cvt.rn.f16x2.f32: one
                  // instruction per pair and direction instead of one per value plus a pack)
                  const float u0 = f[gg * 8 + 2 * i] * s_out, u1 = f[gg * 8 + 2 * i + 1] * s_out;
                  const __half2 h = __floats2half2_rn(u0, u1);
                  const float2 hb = __half22float2(h);
                  hi[i] = h;
                  lo[i] = __floats2half2_rn(u0 - hb.x, u1 - hb.y);
                }
                const size_t off = (((size_t)((c0 >> 3) + gg) * a.Hp + (y + PAD)) * a.Wp + (px + PAD)) * 8;
                *reinterpret_cast<uint4*>(yh + off) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(yh + a.y_plane + off) = *reinterpret_cast<const uint4*>(lo);
              }
            } else if (a.out_mode == 1) {
              float* yf = reinterpret_cast<float*>(a.y);
#pragma unroll
              for (int i = 0; i < 16; i++)
                if (c0 + i < a.Cout) {
                  yf[((size_t)(img * a.y_ctotal + a.y_coff + c0 + i) * a.slice_h + yy) * a.W + px] = f[i];   // (N, C, h, W)
                  amax = fmaxf(amax, fabsf(f[i]));
                }
            } else {  // fused 1x1 head over the (<= 16) channels of this single column group
              float hsum = 0.f;
#pragma unroll
              for (int i = 0; i < 16; i++)
                if (i < a.Cout) hsum = fmaf(__ldg(a.head_w + i), f[i], hsum);
              hsum += __ldg(a.head_b);
              reinterpret_cast<float*>(a.y)[((size_t)img * a.slice_h + yy) * a.W + px] = hsum;                  // (N, 1, h, W)
              amax = fmaxf(amax, fabsf(hsum));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ACC_EMPTY(r));
      }
    }
    if (a.out_meta) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0 && amax > 0.f) atomicMax(&a.out_meta->amax_bits, __float_as_uint(amax));
      if (a.out_mode == 0 && blockIdx.x == 0 && warp == 2 && lane == 0) a.out_meta->scale = s_out;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it / arrive on it
  if (warp == 1) tmem_dealloc(taddr, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// fp32 NCHW -> split chunked planes (channels zero-padded up to Cpad), scale from the measured amax
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_pack_split(const float* __restrict__ x, int C, int Cpad, int H, int W, const ActMeta* __restrict__ in_meta,
                 __half* __restrict__ y, size_t y_plane, ActMeta* __restrict__ out_meta) {
  const int Hp = H + 2 * PAD, Wp = W + 2 * PAD;
  const size_t npix = (size_t)H * W;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float s = pow2_scale_for(__uint_as_float(in_meta->amax_bits));
  if (e == 0) {
    out_meta->scale = s;
    out_meta->amax_bits = in_meta->amax_bits;
  }
  if (e >= npix * (size_t)(Cpad >> 3)) return;
  const size_t pix = e % npix;
  const int j = (int)(e / npix);
  const int py = (int)(pix / W), px = (int)(pix % W);
  __align__(16) __half2 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c0 = j * 8 + 2 * i;
    const float u0 = (c0 < C ? __ldg(x + (size_t)c0 * npix + pix) : 0.f) * s;
    const float u1 = (c0 + 1 < C ? __ldg(x + (size_t)(c0 + 1) * npix + pix) : 0.f) * s;
    const __half h0 = __float2half_rn(u0), h1 = __float2half_rn(u1);
    hi[i] = __halves2half2(h0, h1);
    lo[i] = __halves2half2(__float2half_rn(u0 - __half2float(h0)), __float2half_rn(u1 - __half2float(h1)));
  }
  const size_t off = (((size_t)j * Hp + (py + PAD)) * Wp + (px + PAD)) * 8;
  *reinterpret_cast<uint4*>(y + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(y + y_plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// bilinear sample of F.interpolate(align_corners=False) -- the arithmetic of k_resize_bilinear (conv.cu)
__device__ __forceinline__ float bilinear_at(const float* __restrict__ p, int H, int W, int Ho, int Wo, int j, int i) {
  if (Ho == H && Wo == W) return __ldg(p + (size_t)j * W + i);
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  float fy = sh * ((float)j + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  float fx = sw * ((float)i + 0.5f) - 0.5f;
  fx = fx < 0.f ? 0.f : fx;
  const int yl = (int)fy, xl = (int)fx;
  const int yh = yl + (yl < H - 1 ? 1 : 0), xh = xl + (xl < W - 1 ? 1 : 0);
  const float ly = fy - (float)yl, lx = fx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * __ldg(p + (size_t)yl * W + xl) + lx * __ldg(p + (size_t)yl * W + xh)) +
         ly * (hx * __ldg(p + (size_t)yh * W + xl) + lx * __ldg(p + (size_t)yh * W + xh));
}

// Input of one pyramid level, straight into the split chunked layout (16 channels, the unused ones 0):
//   channels [0, C) = resize(x (C, H, W) -> (h, w)),  channel C = resize(o (1, hc, wc) -> (h, w)) if o
// (multi_scale_net.py:113,117-125: F.upsample + torch.cat).  Replaces 2 resizes + amax + pack.  The
// scale comes from max(amax x, amax o): bilinear interpolation never exceeds the range of its source.
__global__ void __launch_bounds__(256)
    k_pyramid_input(const float* __restrict__ x, int C, int H, int W, const ActMeta* __restrict__ x_meta,
                    const float* __restrict__ o, int hc, int wc, const ActMeta* __restrict__ o_meta, int h, int w,
                    __half* __restrict__ y, size_t y_plane, ActMeta* __restrict__ out_meta, int nimg) {
  // blockIdx.y = image of the batch: its rows start at img * (h + 2 PAD) of the tall split image
  const int img = blockIdx.y;
  x += (size_t)img * C * H * W;
  if (o) o += (size_t)img * hc * wc;
  const int Hp = nimg * (h + 2 * PAD), Wp = w + 2 * PAD;
  const size_t npix = (size_t)h * w;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float bound = __uint_as_float(x_meta->amax_bits);
  if (o) bound = fmaxf(bound, __uint_as_float(o_meta->amax_bits));
  const float s = pow2_scale_for(bound);
  if (e == 0 && img == 0) {
    out_meta->scale = s;
    out_meta->amax_bits = __float_as_uint(bound);
  }
  if (e >= npix) return;
  const int py = (int)(e / w), px = (int)(e % w);
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    v[c] = 0.f;
    if (c < C) v[c] = bilinear_at(x + (size_t)c * H * W, H, W, h, w, py, px);
    else if (c == C && o) v[c] = bilinear_at(o, hc, wc, h, w, py, px);
  }
  __align__(16) __half2 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float u0 = v[2 * i] * s, u1 = v[2 * i + 1] * s;
    const __half h0 = __float2half_rn(u0), h1 = __float2half_rn(u1);
    hi[i] = __halves2half2(h0, h1);
    lo[i] = __halves2half2(__float2half_rn(u0 - __half2float(h0)), __float2half_rn(u1 - __half2float(h1)));
  }
  const size_t off = (((size_t)0 * Hp + (img * (h + 2 * PAD) + py + PAD)) * Wp + (px + PAD)) * 8;   // chunk 0; chunk 1 stays zero
  *reinterpret_cast<uint4*>(y + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(y + y_plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// split chunked planes -> fp32 NCHW (tests / debugging of intermediate layers)
__global__ void __launch_bounds__(256)
    k_unpack_split(const __half* __restrict__ x, size_t x_plane, const ActMeta* __restrict__ meta, int C, int H, int W,
                   float* __restrict__ y) {
  const int Hp = H + 2 * PAD, Wp = W + 2 * PAD;
  const size_t npix = (size_t)H * W;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= npix * (size_t)C) return;
  const size_t pix = e % npix;
  const int c = (int)(e / npix);
  const int py = (int)(pix / W), px = (int)(pix % W);
  const size_t off = (((size_t)(c >> 3) * Hp + (py + PAD)) * Wp + (px + PAD)) * 8 + (c & 7);
  y[e] = (__half2float(x[off]) + __half2float(x[x_plane + off])) / meta->scale;
}

// weight (Cout, Cin, KS, KS) fp32 -> packed split slots [Cin_pad/16][ky][kx][j][plane][Cout_pad][8 halves]
// (zero for the padded input / output channels)
__global__ void __launch_bounds__(256)
    k_pack_weights(const float* __restrict__ w, int Cin, int Cout, int KS, int Cin_pad, int Cout_pad, float w_scale,
                   __half* __restrict__ out) {
  const int taps = KS * KS;
  const size_t total = (size_t)Cin_pad * Cout_pad * taps;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int t = (int)(e % taps), ci = (int)((e / taps) % Cin_pad), n = (int)(e / ((size_t)taps * Cin_pad));
  const int ky = t / KS, kx = t % KS, c = ci >> 4, j = (ci >> 3) & 1, i = ci & 7;
  const float u = (ci < Cin && n < Cout) ? w[((size_t)n * Cin + ci) * taps + t] * w_scale : 0.f;
  const __half h = __float2half_rn(u);
  const __half l = __float2half_rn(u - __half2float(h));
  const size_t slot = ((size_t)(c * KS + ky) * KS + kx) * 2 + j;  // then [plane][n][8]: w_hi rows, then w_lo rows
  out[((slot * 2 + 0) * Cout_pad + n) * 8 + i] = h;
  out[((slot * 2 + 1) * Cout_pad + n) * 8 + i] = l;
}

}  // namespace tc
}  // namespace fnx

// =============================================================================================
// host side
// =============================================================================================
using namespace fnx::tc;

#define FNX_CUDA_TRY(who, call)                                                      \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); \
  } while (0)

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  }
  return n;
}

static size_t act_plane_halves(int C, int H, int W) {
  return (size_t)(C / 8) * (H + 2 * PAD) * (W + 2 * PAD) * 8;
}

static long long* g_tc_debug = nullptr;  // fnx_tc_set_debug

static int pad16(int c) { return (c + 15) / 16 * 16; }
static int cout_pad(int c) { return c <= 16 ? 16 : c <= 32 ? 32 : c <= 64 ? 64 : 128; }

static bool tc_eligible(int Cin, int Cout, int ksize) {
  return (ksize == 3 || ksize == 5) && Cin >= 1 && Cin <= 128 && Cout >= 1 && Cout <= 128 &&
         !(ksize == 5 && Cout > 32);  // 5x5 instantiated for the narrow layers only
}

// Rows per block.  One persistent CTA per SM works through ceil(nblocks / SMs) blocks; a block costs
// max(tensor time, L2->SM fill time) (activation rows incl. halo + the streamed weight slots).  Small
// R wastes L2 bandwidth on halo rows and weight re-streaming, large R leaves an idle tail when the
// block count is not a multiple of the SM count: pick the R in [1, RMAX] with the least modelled time.
template <int KS, int COUT, int RMAX, int WS>
static int choose_rows(const ConvArgs& a) {
  using C = Cfg<KS, COUT, RMAX, WS>;
  const int tiles_x = (a.W + TW - 1) / TW, sms = num_sms();
  const bool resident = a.nchunks * KS <= WS;
  // per tap: MMA1 (N = 2*COUT) + MMA2 (N = COUT); each max(tensor floor N/2, operand read / 128 B per clk)
  const double t_tap = fmax((double)COUT, 32.0 + COUT / 2.0) + fmax(COUT / 2.0, 32.0 + COUT / 4.0);
  const double l2_bytes_per_clk = 24.0;
  if (const char* e = getenv("FNX_TC_ROWS")) {  // tuning override: force R (clamped to the kernel's capacity)
    const int r = atoi(e);
    if (r >= 1) return r < RMAX ? r : RMAX;
  }
  int best = 1;
  double best_t = 1e300;
  for (int r = 1; r <= RMAX; r++) {
    const int nb = tiles_x * ((a.H + r - 1) / r);
    const int rounds = (nb + sms - 1) / sms;
    const double mma = (double)r * KS * KS * a.nchunks * t_tap;
    const double bytes = (double)a.nchunks * ((double)(r + KS - 1) * 4.0 * C::ROWB + (resident ? 0.0 : (double)KS * C::W_STAGE));
    const double t = rounds * (fmax(mma, bytes / l2_bytes_per_clk) + 600.0);
    if (t < best_t * 0.995 || (t <= best_t * 1.005 && r > best)) {
      best_t = t < best_t ? t : best_t;
      best = r;
    }
  }
  return best;
}

template <int KS, int COUT, int RMAX, int WS>
static int launch_tc_rows(ConvArgs a, cudaStream_t st) {
  using C = Cfg<KS, COUT, RMAX, WS>;
  a.R = choose_rows<KS, COUT, RMAX, WS>(a);
  a.tiles_x = (a.W + TW - 1) / TW;
  a.nblocks = a.tiles_x * ((a.H + a.R - 1) / a.R);
  // opt-in (FNX_TC_CLUSTER=1): 2-CTA clusters with multicast weight slots for the layers that stream weights
  static const bool want_cluster = []() { const char* e = getenv("FNX_TC_CLUSTER"); return e && e[0] == '1'; }();
  const bool cluster = want_cluster && a.nchunks * KS > WS && a.nblocks >= 2;
  int grid = a.nblocks < num_sms() ? a.nblocks : num_sms();
  if (cluster) grid = (grid + 1) & ~1;
  if (cluster && grid > num_sms()) grid -= 2;
  a.iters = (a.nblocks + grid - 1) / grid;
  a.dbg = g_tc_debug;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  // PDL is opt-in (FNX_PDL=1): measured on B200 it does not pay -- the next conv CTA cannot co-reside with a
  // running one (shared memory, TMEM), and early-launched grids cost 1-3 % (512^2: 0.633 vs 0.623 ms/step)
  static const bool pdl = []() { const char* e = getenv("FNX_PDL"); return e && e[0] == '1'; }();
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  if (cluster) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    na++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (cluster) {
    auto kern = k_conv_tc<KS, COUT, RMAX, WS, 2>;
    FNX_CUDA_TRY("conv_tc", cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    FNX_CUDA_TRY("conv_tc", cudaLaunchKernelEx(&cfg, kern, a));
  } else {
    auto kern = k_conv_tc<KS, COUT, RMAX, WS, 1>;
    FNX_CUDA_TRY("conv_tc", cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    FNX_CUDA_TRY("conv_tc", cudaLaunchKernelEx(&cfg, kern, a));
  }
  fnx_count_launches(1);
  return FNX_OK;
}

__global__ void __launch_bounds__(256) k_amax(const float* __restrict__ x, size_t n, ActMeta* meta) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&meta->amax_bits, __float_as_uint(m));
}

extern "C" {

size_t fnx_tc_act_bytes(int C, int H, int W) {
  return 2 * act_plane_halves(pad16(C), H, W) * sizeof(__half) + 4096;
}

size_t fnx_tc_weight_bytes(int Cin, int Cout, int ksize) {
  return (size_t)pad16(Cin) * cout_pad(Cout) * ksize * ksize * 2 * sizeof(__half);
}

int fnx_tc_pack_weights(const float* w, int Cin, int Cout, int ksize, float w_scale, void* out, void* stream) {
  if (!tc_eligible(Cin, Cout, ksize))
    return fnx_set_error(FNX_ERR_ARG, "tc_pack_weights: unsupported layer %d->%d k%d", Cin, Cout, ksize);
  const size_t total = (size_t)pad16(Cin) * cout_pad(Cout) * ksize * ksize;
  k_pack_weights<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, Cin, Cout, ksize, pad16(Cin),
                                                                                  cout_pad(Cout), w_scale, (__half*)out);
  fnx_count_launches(1);
  FNX_CUDA_TRY("tc_pack_weights", cudaGetLastError());
  return FNX_OK;
}

int fnx_tc_set_debug(long long* buf) {
  g_tc_debug = buf;  // device buffer of >= 4 * SM-count int64, or NULL to switch the instrumentation off
  return FNX_OK;
}

int fnx_tc_amax(const float* x, size_t n, fnx_act_meta* meta, void* stream) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  k_amax<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, (ActMeta*)meta);
  fnx_count_launches(1);
  FNX_CUDA_TRY("tc_amax", cudaGetLastError());
  return FNX_OK;
}

int fnx_tc_pack_split(const float* x, int C, int H, int W, const fnx_act_meta* in_meta, void* y, fnx_act_meta* out_meta,
                      void* stream) {
  if (C < 1 || H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "tc_pack_split: bad shape");
  const int Cp = pad16(C);
  const size_t total = (size_t)H * W * (Cp / 8);
  k_pack_split<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, C, Cp, H, W, (const ActMeta*)in_meta, (__half*)y, act_plane_halves(Cp, H, W), (ActMeta*)out_meta);
  fnx_count_launches(1);
  FNX_CUDA_TRY("tc_pack_split", cudaGetLastError());
  return FNX_OK;
}

int fnx_tc_unpack_split(const void* x, const fnx_act_meta* meta, int C, int H, int W, float* y, void* stream) {
  if (C < 1 || H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "tc_unpack_split: bad shape");
  const int Cp = pad16(C);
  const size_t total = (size_t)H * W * C;
  k_unpack_split<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, act_plane_halves(Cp, H, W), (const ActMeta*)meta, C, H, W, y);
  fnx_count_launches(1);
  FNX_CUDA_TRY("tc_unpack_split", cudaGetLastError());
  return FNX_OK;
}

static int conv_tc_launch(const void* x, const fnx_act_meta* in_meta, const void* w_packed, const float* bias, int Cin,
                          int Cout, int ksize, int H, int W, int relu, float w_scale, float w_norm, float b_max,
                          int out_mode, void* y, fnx_act_meta* out_meta, int y_channels_total, int y_channel_offset,
                          const float* head_w, const float* head_b, int w_nrep, void* stream, int nimg = 1);

int fnx_conv_tc(const void* x, const fnx_act_meta* in_meta, const void* w_packed, const float* bias, int Cin, int Cout,
                int ksize, int H, int W, int relu, float w_scale, float w_norm, float b_max, int out_mode, void* y,
                fnx_act_meta* out_meta, int y_channels_total, int y_channel_offset, void* stream) {
  if (out_mode != 0 && out_mode != 1) return fnx_set_error(FNX_ERR_ARG, "conv_tc: out_mode must be 0 or 1");
  return conv_tc_launch(x, in_meta, w_packed, bias, Cin, Cout, ksize, H, W, relu, w_scale, w_norm, b_max, out_mode, y,
                        out_meta, y_channels_total, y_channel_offset, nullptr, nullptr, 1, stream);
}

static int conv_tc_launch(const void* x, const fnx_act_meta* in_meta, const void* w_packed, const float* bias, int Cin,
                          int Cout, int ksize, int H, int W, int relu, float w_scale, float w_norm, float b_max,
                          int out_mode, void* y, fnx_act_meta* out_meta, int y_channels_total, int y_channel_offset,
                          const float* head_w, const float* head_b, int w_nrep, void* stream, int nimg) {
  // nimg images of H rows each, stacked as one tall image of nimg * (H + 2 PAD) - 2 PAD rows (see ConvArgs)
  const int slice_h = H;
  H = nimg * (H + 2 * PAD) - 2 * PAD;
  if (!tc_eligible(Cin, Cout, ksize))
    return fnx_set_error(FNX_ERR_ARG, "conv_tc: unsupported layer %d->%d k%d", Cin, Cout, ksize);
  if (H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "conv_tc: bad shape");
  if (out_mode == 0 && (!out_meta || Cout % 16 != 0))
    return fnx_set_error(FNX_ERR_ARG, "conv_tc: split output needs out_meta and Cout %% 16 == 0");
  if (out_mode == 1 && y_channels_total < y_channel_offset + Cout)
    return fnx_set_error(FNX_ERR_ARG, "conv_tc: output channel window out of range");
  if (out_mode == 2 && (Cout > 16 || !head_w || !head_b))
    return fnx_set_error(FNX_ERR_ARG, "conv_tc: the fused 1x1 head needs Cout <= 16 and its weights");
  ConvArgs a;
  a.head_w = head_w; a.head_b = head_b;
  a.w_nrep = w_nrep > 1 ? w_nrep : 1;
  a.w_rep_stride = fnx_tc_weight_bytes(Cin, Cout, ksize);
  a.x = (const __half*)x;
  a.x_plane = act_plane_halves(pad16(Cin), H, W);
  a.w = (const uint8_t*)w_packed;
  a.bias = bias;
  a.y = y;
  a.y_plane = act_plane_halves(pad16(Cout), H, W);
  a.in_meta = (const ActMeta*)in_meta;
  a.out_meta = (ActMeta*)out_meta;
  a.nchunks = pad16(Cin) / 16; a.Cout = Cout;
  a.H = H; a.W = W; a.Hp = H + 2 * PAD; a.Wp = W + 2 * PAD;
  a.relu = relu; a.out_mode = out_mode; a.y_ctotal = y_channels_total; a.y_coff = y_channel_offset;
  a.w_scale = w_scale; a.w_norm = w_norm; a.b_max = b_max;
  a.tiles_x = 0; a.nblocks = 0; a.R = 1; a.iters = 0; a.dbg = nullptr;
  a.slice_h = slice_h; a.slice_pitch = slice_h + 2 * PAD;
  cudaStream_t st = (cudaStream_t)stream;
  const int cp = cout_pad(Cout);
  if (ksize == 3) {
    switch (cp) {
      case 128: return launch_tc_rows<3, 128, 2, 5>(a, st);  // 5 weight slots leave room for a 3rd activation stage
      case 64: return launch_tc_rows<3, 64, 4, 6>(a, st);
      case 32: return launch_tc_rows<3, 32, 8, 6>(a, st);
      default: return launch_tc_rows<3, 16, 8, 6>(a, st);
    }
  }
  // 5x5: weight rings sized so the net's two 5x5 layers (1 and 2 chunks) keep their weights resident
  return cp == 32 ? launch_tc_rows<5, 32, 4, 6>(a, st) : launch_tc_rows<5, 16, 4, 10>(a, st);
}

// ---------------------------------------------------------------------------------------------
// MultiScaleNet.forward (multi_scale_net.py:101-127): one call, every launch enqueued on `stream`
// ---------------------------------------------------------------------------------------------
// ---- optional per-layer timing (bench.py roofline): CUDA events around every conv launch -------
namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> pool;   // pairs (begin, end)
  std::vector<fnx_profile_rec> recs;
  std::mutex mu;
} g_prof;

struct LayerTimer {
  cudaStream_t st;
  bool active = false;
  size_t idx = 0;
  LayerTimer(const fnx_conv_layer& l, int h, int w, int tensor, cudaStream_t s) : st(s) {
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    idx = g_prof.recs.size();
    while (g_prof.pool.size() < 2 * (idx + 1)) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      g_prof.pool.push_back(e);
    }
    g_prof.recs.push_back(fnx_profile_rec{l.cin, l.cout, l.ksize, h, w, tensor, 0.f});
    cudaEventRecord(g_prof.pool[2 * idx], st);
    active = true;
  }
  ~LayerTimer() {
    if (active) cudaEventRecord(g_prof.pool[2 * idx + 1], st);
  }
};
}  // namespace

int fnx_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.on = enable != 0;
  g_prof.recs.clear();
  return FNX_OK;
}

int fnx_profile_fetch(fnx_profile_rec* out, int capacity) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  const int n = (int)g_prof.recs.size();
  for (int i = 0; i < n; i++) {
    if (cudaEventSynchronize(g_prof.pool[2 * i + 1]) != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "profile_fetch: event sync failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]);
    g_prof.recs[i].ms = ms;
    if (out && i < capacity) out[i] = g_prof.recs[i];
  }
  g_prof.recs.clear();
  return n;
}

namespace {
struct Bump {
  uint8_t* base;
  size_t off = 0;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  }
};
constexpr int MAX_META = 64;

struct Runner {
  const fnx_msnet_plan* plan;
  Bump ws;
  bool dry;
  cudaStream_t st;
  int nimg = 1;   // images per forward (tensor-core plan: the whole batch in one launch per layer, as a tall image)
  int tall(int h) const { return nimg * (h + 2 * PAD) - 2 * PAD; }
  int nmeta = 0;
  ActMeta* metas;
  ActMeta* new_meta() { return metas ? metas + (nmeta++ % MAX_META) : (nmeta++, nullptr); }

  static bool is_tc(const fnx_conv_layer& l) { return l.w_tc && tc_eligible(l.cin, l.cout, l.ksize); }

  // runs one Sequential of convs at resolution (h, w).  Input: `in` fp32 NCHW with layers[0].cin channels,
  // or (in_split, in_meta) already in the split chunked layout.  The last layer's output goes to `out`
  // (fp32 NCHW, out_ctotal channels, window at out_coff); `last_meta` (optional) receives its max|.|;
  // `head` (optional) = a 1x1 conv fused into the last layer's epilogue, `out` is then its 1-channel result.
  int block(const fnx_conv_layer* L, int n, const float* in, const void* in_split, ActMeta* in_meta, bool split_in,
            int h, int w, float* out, int out_ctotal, int out_coff, ActMeta* last_meta = nullptr,
            const fnx_conv_layer* head = nullptr) {
    const float* cur_f32 = in;
    const void* cur_split = in_split;
    bool is_split = split_in;  // which of the two holds the current tensor (pointers are null in a dry run)
    ActMeta* cur_meta = in_meta;         // amax valid (and scale, when split)
    for (int i = 0; i < n; i++) {
      const fnx_conv_layer& l = L[i];
      const bool last = i == n - 1;
      const bool tc_now = is_tc(l);
      const bool tc_next = !last && is_tc(L[i + 1]);
      if (tc_now) {
        if (!is_split) {
          if (!cur_meta) {
            cur_meta = new_meta();
            if (!dry) { int rc = fnx_tc_amax(cur_f32, (size_t)l.cin * h * w, (fnx_act_meta*)cur_meta, st); if (rc) return rc; }
          }
          if (nimg != 1) return fnx_set_error(FNX_ERR_ARG, "msnet: a batched forward needs the all-tensor-core plan");
          void* sp = ws.take(fnx_tc_act_bytes(l.cin, h, w));
          ActMeta* m = new_meta();
          if (!dry) { int rc = fnx_tc_pack_split(cur_f32, l.cin, h, w, (fnx_act_meta*)cur_meta, sp, (fnx_act_meta*)m, st); if (rc) return rc; }
          cur_split = sp; cur_meta = m; is_split = true;
        }
        if (tc_next) {
          void* sp = ws.take(fnx_tc_act_bytes(l.cout, tall(h), w));
          ActMeta* m = new_meta();
          if (!dry) {
            LayerTimer lt(l, h * nimg, w, 1, st);   // (a batched launch: the rows of all its images)
            int rc = conv_tc_launch(cur_split, (fnx_act_meta*)cur_meta, l.w_tc, l.bias, l.cin, l.cout, l.ksize, h, w, l.relu,
                                    l.w_scale, l.w_norm, l.b_max, 0, sp, (fnx_act_meta*)m, 0, 0, nullptr, nullptr,
                                    l.w_replicas, st, nimg);
            if (rc) return rc;
          }
          cur_split = sp; cur_meta = m; cur_f32 = nullptr; is_split = true;
        } else {
          const bool fuse_head = last && head != nullptr;
          float* o = last ? out : (float*)ws.take((size_t)nimg * l.cout * h * w * 4);
          if (!dry) {
            LayerTimer lt(l, h * nimg, w, 1, st);
            int rc = conv_tc_launch(cur_split, (fnx_act_meta*)cur_meta, l.w_tc, l.bias, l.cin, l.cout, l.ksize, h, w, l.relu,
                                    l.w_scale, l.w_norm, l.b_max, fuse_head ? 2 : 1, o,
                                    (fnx_act_meta*)(last ? last_meta : nullptr), last ? out_ctotal : l.cout,
                                    last ? out_coff : 0, fuse_head ? head->weight : nullptr,
                                    fuse_head ? head->bias : nullptr, l.w_replicas, st, nimg);
            if (rc) return rc;
          }
          cur_f32 = o; cur_split = nullptr; cur_meta = nullptr; is_split = false;
        }
      } else {
        if (is_split) return fnx_set_error(FNX_ERR_ARG, "msnet: internal layout mismatch");
        if (nimg != 1) return fnx_set_error(FNX_ERR_ARG, "msnet: a batched forward needs the all-tensor-core plan");
        float* o = last ? out : (float*)ws.take((size_t)l.cout * h * w * 4);
        ActMeta* m = tc_next ? new_meta() : (last ? last_meta : nullptr);
        if (!dry) {
          LayerTimer lt(l, h, w, 0, st);
          int rc = fnx_conv_direct(cur_f32, l.weight, l.bias, o, 1, l.cin, h, w, l.cout, l.ksize, l.relu,
                                   last ? out_ctotal : l.cout, last ? out_coff : 0, m ? &m->amax_bits : nullptr, st);
          if (rc) return rc;
        }
        cur_f32 = o; cur_split = nullptr; cur_meta = tc_next ? m : nullptr; is_split = false;
      }
    }
    return FNX_OK;
  }

  // input of one pyramid level in the split chunked layout (k_pyramid_input)
  int pyramid_input(const float* x, int c, int H, int W, ActMeta* x_meta, const float* o, int hc, int wc, ActMeta* o_meta,
                    int h, int w, void** split_out, ActMeta** meta_out) {
    void* sp = ws.take(fnx_tc_act_bytes(16, tall(h), w));
    ActMeta* m = new_meta();
    if (!dry) {
      const size_t npix = (size_t)h * w;
      k_pyramid_input<<<dim3((unsigned)((npix + 255) / 256), (unsigned)nimg), 256, 0, st>>>(
          x, c, H, W, x_meta, o, hc, wc, o_meta, h, w, (__half*)sp, act_plane_halves(16, tall(h), w), m, nimg);
      fnx_count_launches(1);
      FNX_CUDA_TRY("msnet", cudaGetLastError());
    }
    *split_out = sp; *meta_out = m;
    return FNX_OK;
  }

  int forward_one(const float* x, float* y, int H, int W) {
    const int c = plan->data_channels;
    const int h4 = (int)(H * 0.25), w4 = (int)(W * 0.25), h2 = (int)(H * 0.5), w2 = (int)(W * 0.5);
    if (h4 < 1 || w4 < 1) return fnx_set_error(FNX_ERR_ARG, "msnet: grid %dx%d too small for the 1/4 scale", H, W);
    metas = (ActMeta*)ws.take(MAX_META * sizeof(ActMeta));
    if (!dry) FNX_CUDA_TRY("msnet", cudaMemsetAsync(metas, 0, MAX_META * sizeof(ActMeta), st));
    int rc;
    // the 1x1 `final` conv rides in the epilogue of the last full-resolution layer when that one is a
    // tensor-core layer with <= 16 output channels
    const fnx_conv_layer& lastf = plan->full[5];
    const fnx_conv_layer& fin = plan->final_conv;
    const bool fuse_head = is_tc(lastf) && lastf.cout <= 16 && fin.ksize == 1 && fin.cin == lastf.cout && fin.cout == 1 &&
                           !fin.relu;
    float* o4 = (float*)ws.take((size_t)nimg * h4 * w4 * 4);
    float* o2 = (float*)ws.take((size_t)nimg * h2 * w2 * 4);
    float* o1 = fuse_head ? nullptr : (float*)ws.take((size_t)nimg * lastf.cout * H * W * 4);
    if (is_tc(plan->quarter[0]) && is_tc(plan->half[0]) && is_tc(plan->full[0]) && c + 1 <= 8) {
      // tensor-core plan: every level's input goes straight into the split layout
      ActMeta *mx = new_meta(), *mo4 = new_meta(), *mo2 = new_meta(), *m;
      void* sp;
      if (!dry && (rc = fnx_tc_amax(x, (size_t)nimg * c * H * W, (fnx_act_meta*)mx, st))) return rc;
      if ((rc = pyramid_input(x, c, H, W, mx, nullptr, 0, 0, nullptr, h4, w4, &sp, &m))) return rc;
      if ((rc = block(plan->quarter, 4, nullptr, sp, m, true, h4, w4, o4, 1, 0, mo4))) return rc;
      if ((rc = pyramid_input(x, c, H, W, mx, o4, h4, w4, mo4, h2, w2, &sp, &m))) return rc;
      if ((rc = block(plan->half, 6, nullptr, sp, m, true, h2, w2, o2, 1, 0, mo2))) return rc;
      if ((rc = pyramid_input(x, c, H, W, mx, o2, h2, w2, mo2, H, W, &sp, &m))) return rc;
      if (fuse_head) return block(plan->full, 6, nullptr, sp, m, true, H, W, y, 1, 0, nullptr, &fin);
      if (nimg != 1) return fnx_set_error(FNX_ERR_ARG, "msnet: a batched forward needs the fused 1x1 head");
      if ((rc = block(plan->full, 6, nullptr, sp, m, true, H, W, o1, lastf.cout, 0))) return rc;
      return block(&fin, 1, o1, nullptr, nullptr, false, H, W, y, 1, 0);
    }
    if (nimg != 1) return fnx_set_error(FNX_ERR_ARG, "msnet: a batched forward needs the all-tensor-core plan");
    float* x4 = (float*)ws.take((size_t)c * h4 * w4 * 4);
    float* in2 = (float*)ws.take((size_t)(c + 1) * h2 * w2 * 4);
    float* in1 = (float*)ws.take((size_t)(c + 1) * H * W * 4);
#define RUN(expr) do { if (!dry) { rc = (expr); if (rc) return rc; } } while (0)
    RUN(fnx_resize_bilinear(x, x4, 1, c, H, W, h4, w4, c, 0, st));
    if ((rc = block(plan->quarter, 4, x4, nullptr, nullptr, false, h4, w4, o4, 1, 0))) return rc;
    RUN(fnx_resize_bilinear(x, in2, 1, c, H, W, h2, w2, c + 1, 0, st));
    RUN(fnx_resize_bilinear(o4, in2, 1, 1, h4, w4, h2, w2, c + 1, c, st));
    if ((rc = block(plan->half, 6, in2, nullptr, nullptr, false, h2, w2, o2, 1, 0))) return rc;
    RUN(fnx_resize_bilinear(x, in1, 1, c, H, W, H, W, c + 1, 0, st));
    RUN(fnx_resize_bilinear(o2, in1, 1, 1, h2, w2, H, W, c + 1, c, st));
#undef RUN
    if (fuse_head) return block(plan->full, 6, in1, nullptr, nullptr, false, H, W, y, 1, 0, nullptr, &fin);
    if ((rc = block(plan->full, 6, in1, nullptr, nullptr, false, H, W, o1, lastf.cout, 0))) return rc;
    return block(&fin, 1, o1, nullptr, nullptr, false, H, W, y, 1, 0);
  }
};
}  // namespace

// can the whole batch go through one launch per layer?  (the plan a shipped ScaleNet gets: every layer on the
// tensor cores, pyramid inputs written straight in the split layout, the 1x1 head fused)
static bool batchable(const fnx_msnet_plan* plan) {
  const fnx_conv_layer& lastf = plan->full[5];
  const fnx_conv_layer& fin = plan->final_conv;
  auto tc = [](const fnx_conv_layer& l) { return l.w_tc && tc_eligible(l.cin, l.cout, l.ksize); };
  for (int i = 0; i < 4; i++) if (!tc(plan->quarter[i])) return false;
  for (int i = 0; i < 6; i++) if (!tc(plan->half[i]) || !tc(plan->full[i])) return false;
  return plan->data_channels + 1 <= 8 && lastf.cout <= 16 && fin.ksize == 1 && fin.cin == lastf.cout && fin.cout == 1 && !fin.relu;
}

// Images per launch of a batched forward: the largest divisor of N up to FNX_MSNET_GROUP_MAX.  (Every group must
// have the same size: the chunk stride of the split layout -- hence where the zero pad rows live -- depends on it.
// 72 images of 256^2 keep the workspace near 8 GB; a group is already one launch per layer for 72 images.)
#ifndef FNX_MSNET_GROUP_MAX
#define FNX_MSNET_GROUP_MAX 72
#endif
static int batch_group(const fnx_msnet_plan* plan, int N) {
  if (N <= 1 || !batchable(plan)) return 1;
  int g = 1;
  for (int d = 1; d <= N && d <= FNX_MSNET_GROUP_MAX; d++)
    if (N % d == 0) g = d;
  return g;
}

size_t fnx_msnet_workspace(const fnx_msnet_plan* plan, int H, int W) { return fnx_msnet_workspace_n(plan, 1, H, W); }

size_t fnx_msnet_workspace_n(const fnx_msnet_plan* plan, int N, int H, int W) {
  if (!plan || N < 1) return 0;
  Runner r{plan, Bump{nullptr}, true, nullptr};
  r.nimg = batch_group(plan, N);
  if (r.forward_one(nullptr, nullptr, H, W) != FNX_OK) return 0;
  return r.ws.off + 4096;
}

int fnx_msnet_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  FNX_CUDA_TRY("msnet_workspace_init", cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));
  return FNX_OK;
}

int fnx_msnet_forward(const fnx_msnet_plan* plan, const float* x, float* y, int N, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (!plan || N < 1 || H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "msnet_forward: bad arguments");
  const size_t need = fnx_msnet_workspace_n(plan, N, H, W);
  if (need == 0) return FNX_ERR_ARG;
  if (!workspace || workspace_bytes < need)
    return fnx_set_error(FNX_ERR_WORKSPACE, "msnet_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  const int c = plan->data_channels;
  const int group = batch_group(plan, N);
  if (group > 1) {
    // groups of images as one tall image each: one launch per layer and group (activation scales shared by a group)
    for (int n = 0; n < N; n += group) {
      Runner r{plan, Bump{(uint8_t*)workspace}, false, (cudaStream_t)stream};
      r.nimg = group;
      int rc = r.forward_one(x + (size_t)n * c * H * W, y + (size_t)n * H * W, H, W);
      if (rc) return rc;
    }
    return FNX_OK;
  }
  for (int n = 0; n < N; n++) {
    Runner r{plan, Bump{(uint8_t*)workspace}, false, (cudaStream_t)stream};
    int rc = r.forward_one(x + (size_t)n * c * H * W, y + (size_t)n * H * W, H, W);
    if (rc) return rc;
  }
  return FNX_OK;
}

}  // extern "C"
