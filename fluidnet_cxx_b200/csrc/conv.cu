// conv.cu -- FluidNet / MultiScaleNet forward pieces (sm_100a): fp32 direct convolution with
// fused bias + ReLU, bilinear resize (align_corners = False), and the wrapper's input/output
// stencils.  Reference: pytorch/lib/multi_scale_net.py:101-127 (17 Conv2d, zero padding k//2,
// F.upsample bilinear), trained_models/.../*_saved.py:78-238 (FluidNet.forward: div -> std
// normalise -> [div/s, occupancy] -> net -> velocityUpdate -> un-normalise -> setWallBcs).
//
// This translation unit is compiled WITH FMA contraction (it is a dense contraction whose
// reference arithmetic lives in PyTorch's conv kernels; parity bar 1e-5 relative).
// The tcgen05 implicit-GEMM path of every 3x3 / 5x5 layer lives in conv_tc.cu; this file is the fp32
// direct convolution (any kernel size; the numerical cross-check of the tensor-core path and the
// `USE_TENSOR_CORES = False` plan), the standalone resize and the FluidNet wrapper stencils.
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"
#include "fluid_common.cuh"
#include "cnn_internal.h"
#include "host_util.h"
#include "stencil_device.cuh"

namespace fnx {

// ---------------------------------------------------------------------------------------------
// direct convolution, NCHW fp32, stride 1, zero padding KS/2
//   block = 128 threads: tx 0..7 (x phase), ty 0..3 (row), tz 0..3 (output-channel group)
//   tile  = 64 x 4 pixels x (4*COT) output channels; thread = 8 pixels (x = tx + 8*q) x COT channels
//   input channels are staged through shared memory CIC at a time
// ---------------------------------------------------------------------------------------------
constexpr int CV_TW = 64, CV_TH = 4, CV_CIC = 8;

template <int KS, int COT>
__global__ void __launch_bounds__(128)
    k_conv_direct(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int Cin, int Cout, int H, int W, int relu, int y_ctotal, int y_coff,
                  unsigned* __restrict__ amax_bits) {
  constexpr int PAD = KS / 2;
  constexpr int SW = CV_TW + KS - 1, SH = CV_TH + KS - 1;
  constexpr int CO_BLK = 4 * COT;
  __shared__ float s_in[CV_CIC][SH][SW + 1];
  __shared__ __align__(16) float s_w[CV_CIC][KS * KS][CO_BLK];
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = (tid >> 3) & 3, tz = tid >> 5;
  const int x0 = blockIdx.x * CV_TW, y0 = blockIdx.y * CV_TH;
  const int nco_blk = (Cout + CO_BLK - 1) / CO_BLK;
  const int n = blockIdx.z / nco_blk, co0 = (blockIdx.z % nco_blk) * CO_BLK;
  x += (size_t)n * Cin * H * W;
  y += (size_t)n * y_ctotal * H * W;

  float acc[8][COT];
#pragma unroll
  for (int q = 0; q < 8; q++)
#pragma unroll
    for (int c = 0; c < COT; c++) acc[q][c] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CV_CIC) {
    // stage the input tile (zero padded) and the weight slice
    for (int e = tid; e < CV_CIC * SH * SW; e += 128) {
      const int cc = e / (SH * SW), r = (e / SW) % SH, col = e % SW;
      const int gy = y0 + r - PAD, gx = x0 + col - PAD, ci = ci0 + cc;
      float v = 0.f;
      if (ci < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(x + ((size_t)ci * H + gy) * W + gx);
      s_in[cc][r][col] = v;
    }
    for (int e = tid; e < CV_CIC * KS * KS * CO_BLK; e += 128) {
      const int cc = e / (KS * KS * CO_BLK), t = (e / CO_BLK) % (KS * KS), c = e % CO_BLK;
      const int ci = ci0 + cc, co = co0 + c;
      float v = 0.f;
      if (ci < Cin && co < Cout) v = __ldg(w + ((size_t)co * Cin + ci) * KS * KS + t);
      s_w[cc][t][c] = v;
    }
    __syncthreads();
    const int cmax = (Cin - ci0) < CV_CIC ? (Cin - ci0) : CV_CIC;
    for (int cc = 0; cc < cmax; cc++) {
#pragma unroll
      for (int dy = 0; dy < KS; dy++) {
#pragma unroll
        for (int dx = 0; dx < KS; dx++) {
          float wv[COT];
#pragma unroll
          for (int c = 0; c < COT; c++) wv[c] = s_w[cc][dy * KS + dx][tz * COT + c];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float v = s_in[cc][ty + dy][tx + 8 * q + dx];
#pragma unroll
            for (int c = 0; c < COT; c++) acc[q][c] = fmaf(v, wv[c], acc[q][c]);
          }
        }
      }
    }
    __syncthreads();
  }
  const int gy = y0 + ty;
  float amax = 0.f;
#pragma unroll
  for (int c = 0; c < COT; c++) {
    const int co = co0 + tz * COT + c;
    if (co >= Cout || gy >= H) continue;
    const float b = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int gx = x0 + tx + 8 * q;
      if (gx >= W) continue;
      float v = acc[q][c] + b;
      if (relu) v = v > 0.f ? v : 0.f;
      amax = fmaxf(amax, fabsf(v));
      y[((size_t)(y_coff + co) * H + gy) * W + gx] = v;
    }
  }
  if (amax_bits) {  // max |y| of the layer, consumed by the split-fp16 packer of the tensor path
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((tid & 31) == 0 && amax > 0.f) atomicMax(amax_bits, __float_as_uint(amax));
  }
}

// ---------------------------------------------------------------------------------------------
// F.interpolate(mode='bilinear', align_corners=False) (Q18), written into channels
// [y_coff, y_coff+C) of a (N, y_ctotal, Ho, Wo) tensor so torch.cat needs no extra pass
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_resize_bilinear(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W, int Ho, int Wo,
                      int y_ctotal, int y_coff) {
  const size_t total = (size_t)N * C * Ho * Wo;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int i = e % Wo, j = (e / Wo) % Ho, c = (e / ((size_t)Wo * Ho)) % C, n = e / ((size_t)Wo * Ho * C);
  const float* p = x + ((size_t)n * C + c) * H * W;
  float v;
  if (Ho == H && Wo == W) {
    v = __ldg(p + (size_t)j * W + i);  // same size: identity
  } else {
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
    float fy = sh * ((float)j + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    float fx = sw * ((float)i + 0.5f) - 0.5f;
    fx = fx < 0.f ? 0.f : fx;
    const int yl = (int)fy, xl = (int)fx;
    const int yh = yl + (yl < H - 1 ? 1 : 0), xh = xl + (xl < W - 1 ? 1 : 0);
    const float ly = fy - (float)yl, lx = fx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
    v = hy * (hx * __ldg(p + (size_t)yl * W + xl) + lx * __ldg(p + (size_t)yl * W + xh)) +
        ly * (hx * __ldg(p + (size_t)yh * W + xl) + lx * __ldg(p + (size_t)yh * W + xh));
  }
  y[(((size_t)n * y_ctotal + y_coff + c) * Ho + j) * Wo + i] = v;
}

// ---------------------------------------------------------------------------------------------
// FluidNet wrapper stencils (2-D): net input and post-processing
// ---------------------------------------------------------------------------------------------
// x[:,0] = velocityDivergence(U, flags) / s ; x[:,1] = flagsToOccupancy(flags)   (*_saved.py:135-177)
__global__ void __launch_bounds__(256)
    k_cnn_input(const float* __restrict__ U, const float* __restrict__ flags, const float* __restrict__ scale,
                float* __restrict__ x, int B, int H, int W) {
  const size_t n = (size_t)H * W;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)B * n) return;
  const int b = e / n;
  const size_t o = e % n;
  const int j = o / W, i = o % W;
  const float* u = U + (size_t)b * 2 * n;
  const float f = __ldg(flags + (size_t)b * n + o);
  float d = 0.f;
  if (!(i < 1 || i > W - 2 || j < 1 || j > H - 2)) {
    d = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(__ldg(u + o), __ldg(u + o + 1)), __ldg(u + n + o)), __ldg(u + n + o + W)), 0.f);
  }
  if (f == kObstacle) d = 0.f;
  x[(size_t)b * 2 * n + o] = __fdiv_rn(d, __ldg(scale + b));
  x[(size_t)b * 2 * n + n + o] = occupancy_of(f);
}

// U/s -> velocityUpdate(p) -> *s -> setWallBcs ; p_out = p*s                      (*_saved.py:149-232)
__global__ void __launch_bounds__(256)
    k_cnn_output(const float* __restrict__ pnet, const float* __restrict__ U, const float* __restrict__ flags,
                 const float* __restrict__ scale, float* __restrict__ p_out, float* __restrict__ U_out, int B, int H,
                 int W, int wall_bcs) {
  const size_t n = (size_t)H * W;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)B * n) return;
  const int b = e / n;
  const size_t o = e % n;
  const int j = o / W, i = o % W;
  const float s = __ldg(scale + b);
  const float* u = U + (size_t)b * 2 * n;
  const float* fl = flags + (size_t)b * n;
  const float* pn = pnet + (size_t)b * n;
  const float fc = __ldg(fl + o), P = __ldg(pn + o);
  const bool interior = !(i < 1 || i > W - 2 || j < 1 || j > H - 2);
  const int idx[2] = {i, j};
  const size_t nb[2] = {1, (size_t)W};
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const float fn = idx[c] > 0 ? __ldg(fl + o - nb[c]) : fc;
    float v = __fdiv_rn(__ldg(u + c * n + o), s);
    if (interior) {
      // velocity_update_apply without FMA contraction (this TU allows contraction elsewhere)
      const float Pn = __ldg(pn + o - nb[c]);
      const bool cf = fc == kFluid, ce = fc == kEmpty;
      const float m1 = (cf && fn == kFluid) ? 1.f : 0.f, m2 = (cf && fn == kEmpty) ? 1.f : 0.f;
      const float m3 = (ce && fn == kFluid) ? 1.f : 0.f;
      const float t1 = __fmul_rn(m1, __fsub_rn(v, __fsub_rn(P, Pn)));
      const float t2 = __fmul_rn(m2, __fsub_rn(v, P));
      const float t3 = __fmul_rn(m3, __fadd_rn(v, Pn));
      v = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), 0.f);
    }
    v = __fmul_rn(v, s);
    if (wall_bcs) v = wall_bcs_apply(v, fc, fn);
    U_out[(size_t)b * 2 * n + c * n + o] = v;
  }
  p_out[(size_t)b * n + o] = __fmul_rn(P, s);
}

// ---- 3-D, slice-wise projection (FluidNet.forward_fields_3d: this package's definition, no reference counterpart) ----
// x[(b*D + k), 0] = velocityDivergence_3D(U, flags) / s ; x[(b*D + k), 1] = flagsToOccupancy(flags): one image per z-slice
__global__ void __launch_bounds__(256)
    k_cnn_input_3d(const float* __restrict__ U, const float* __restrict__ flags, const float* __restrict__ scale,
                   float* __restrict__ x, int B, int D, int H, int W) {
  const size_t hw = (size_t)H * W, n = (size_t)D * hw;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)B * n) return;
  const int b = e / n;
  const size_t o = e % n;
  const int k = o / hw;
  const size_t o2 = o % hw;
  const int j = o2 / W, i = o2 % W;
  const float* u = U + (size_t)b * 3 * n;
  const float f = __ldg(flags + (size_t)b * n + o);
  float d = 0.f;
  if (!(i < 1 || i > W - 2 || j < 1 || j > H - 2 || k < 1 || k > D - 2)) {
    // velocity_divergence.py:61-69 order: ((u_i - u_{i+1}) + v_j - v_{j+1}) + (w_k - w_{k+1})
    d = __fsub_rn(__fadd_rn(__fsub_rn(__ldg(u + o), __ldg(u + o + 1)), __ldg(u + n + o)), __ldg(u + n + o + W));
    d = __fadd_rn(d, __fsub_rn(__ldg(u + 2 * n + o), __ldg(u + 2 * n + o + hw)));
  }
  if (f == kObstacle) d = 0.f;
  float* xi = x + ((size_t)b * D + k) * 2 * hw;
  xi[o2] = __fdiv_rn(d, __ldg(scale + b));
  xi[hw + o2] = occupancy_of(f);
}

// U/s -> in-plane velocityUpdate(p) on (Ux, Uy), Uz kept -> *s -> setWallBcs_3D ; p_out = p*s
__global__ void __launch_bounds__(256)
    k_cnn_output_3d(const float* __restrict__ pnet, const float* __restrict__ U, const float* __restrict__ flags,
                    const float* __restrict__ scale, float* __restrict__ p_out, float* __restrict__ U_out, int B, int D,
                    int H, int W) {
  const size_t hw = (size_t)H * W, n = (size_t)D * hw;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)B * n) return;
  const int b = e / n;
  const size_t o = e % n;
  const int k = o / hw;
  const size_t o2 = o % hw;
  const int j = o2 / W, i = o2 % W;
  const float s = __ldg(scale + b);
  const float* u = U + (size_t)b * 3 * n;
  const float* fl = flags + (size_t)b * n;
  const float* pn = pnet + (size_t)b * n;      // (B*D, 1, H, W) images in slice order == (B, 1, D, H, W)
  const float fc = __ldg(fl + o), P = __ldg(pn + o);
  const bool interior = !(i < 1 || i > W - 2 || j < 1 || j > H - 2 || k < 1 || k > D - 2);
  const int idx[3] = {i, j, k};
  const size_t nb[3] = {1, (size_t)W, hw};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float fn = idx[c] > 0 ? __ldg(fl + o - nb[c]) : fc;
    float v = __fdiv_rn(__ldg(u + c * n + o), s);
    if (c < 2 && interior) {
      const float Pn = __ldg(pn + o - nb[c]);
      const bool cf = fc == kFluid, ce = fc == kEmpty;
      const float m1 = (cf && fn == kFluid) ? 1.f : 0.f, m2 = (cf && fn == kEmpty) ? 1.f : 0.f;
      const float m3 = (ce && fn == kFluid) ? 1.f : 0.f;
      const float t1 = __fmul_rn(m1, __fsub_rn(v, __fsub_rn(P, Pn)));
      const float t2 = __fmul_rn(m2, __fsub_rn(v, P));
      const float t3 = __fmul_rn(m3, __fadd_rn(v, Pn));
      v = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), 0.f);
    }
    v = __fmul_rn(v, s);
    v = wall_bcs_apply(v, fc, fn);
    U_out[(size_t)b * 3 * n + c * n + o] = v;
  }
  p_out[(size_t)b * n + o] = __fmul_rn(P, s);
}

}  // namespace fnx

using namespace fnx;

#define FNX_CUDA_TRY(who, call)                                                      \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); \
  } while (0)

template <int KS>
static int launch_conv(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int Cout, int H,
                       int W, int relu, int y_ctotal, int y_coff, unsigned* amax, cudaStream_t st) {
  dim3 block(128);
  if (Cout >= 8) {
    constexpr int COT = 8;
    dim3 grid((W + CV_TW - 1) / CV_TW, (H + CV_TH - 1) / CV_TH, N * ((Cout + 4 * COT - 1) / (4 * COT)));
    k_conv_direct<KS, COT><<<grid, block, 0, st>>>(x, w, bias, y, Cin, Cout, H, W, relu, y_ctotal, y_coff, amax);
  } else {
    constexpr int COT = 1;
    dim3 grid((W + CV_TW - 1) / CV_TW, (H + CV_TH - 1) / CV_TH, N * ((Cout + 4 * COT - 1) / (4 * COT)));
    k_conv_direct<KS, COT><<<grid, block, 0, st>>>(x, w, bias, y, Cin, Cout, H, W, relu, y_ctotal, y_coff, amax);
  }
  fnx_count_launches(1);
  FNX_CUDA_TRY("conv2d", cudaGetLastError());
  return FNX_OK;
}

int fnx_conv_direct(const float* x, const float* weight, const float* bias, float* y, int N, int Cin, int H, int W,
                    int Cout, int ksize, int relu, int y_channels_total, int y_channel_offset, unsigned* amax_bits,
                    cudaStream_t st) {
  if (N < 1 || Cin < 1 || Cout < 1 || H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "conv2d: bad shape");
  if (y_channels_total < y_channel_offset + Cout) return fnx_set_error(FNX_ERR_ARG, "conv2d: output channel window out of range");
  switch (ksize) {
    case 1: return launch_conv<1>(x, weight, bias, y, N, Cin, Cout, H, W, relu, y_channels_total, y_channel_offset, amax_bits, st);
    case 3: return launch_conv<3>(x, weight, bias, y, N, Cin, Cout, H, W, relu, y_channels_total, y_channel_offset, amax_bits, st);
    case 5: return launch_conv<5>(x, weight, bias, y, N, Cin, Cout, H, W, relu, y_channels_total, y_channel_offset, amax_bits, st);
    default: return fnx_set_error(FNX_ERR_ARG, "conv2d: kernel size %d not supported (1, 3, 5)", ksize);
  }
}

extern "C" {

int fnx_conv2d(const float* x, const float* weight, const float* bias, float* y, int N, int Cin, int H, int W,
               int Cout, int ksize, int relu, int y_channels_total, int y_channel_offset, void* stream) {
  return fnx_conv_direct(x, weight, bias, y, N, Cin, H, W, Cout, ksize, relu, y_channels_total, y_channel_offset,
                         nullptr, (cudaStream_t)stream);
}

int fnx_resize_bilinear(const float* x, float* y, int N, int C, int H, int W, int Ho, int Wo, int y_channels_total,
                        int y_channel_offset, void* stream) {
  if (N < 1 || C < 1 || H < 1 || W < 1 || Ho < 1 || Wo < 1) return fnx_set_error(FNX_ERR_ARG, "resize_bilinear: bad shape");
  if (y_channels_total < y_channel_offset + C) return fnx_set_error(FNX_ERR_ARG, "resize_bilinear: output channel window out of range");
  const size_t total = (size_t)N * C * Ho * Wo;
  k_resize_bilinear<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, C, H, W, Ho, Wo,
                                                                                      y_channels_total, y_channel_offset);
  fnx_count_launches(1);
  FNX_CUDA_TRY("resize_bilinear", cudaGetLastError());
  return FNX_OK;
}

int fnx_fluidnet_input(const float* U, const float* flags, const float* scale, float* x, int B, int H, int W,
                       void* stream) {
  if (B < 1 || H < 2 || W < 2) return fnx_set_error(FNX_ERR_ARG, "fluidnet_input: bad shape");
  const size_t total = (size_t)B * H * W;
  k_cnn_input<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(U, flags, scale, x, B, H, W);
  fnx_count_launches(1);
  FNX_CUDA_TRY("fluidnet_input", cudaGetLastError());
  return FNX_OK;
}

int fnx_fluidnet_output(const float* p_net, const float* U, const float* flags, const float* scale, float* p_out,
                        float* U_out, int B, int H, int W, int apply_wall_bcs, void* stream) {
  if (B < 1 || H < 2 || W < 2) return fnx_set_error(FNX_ERR_ARG, "fluidnet_output: bad shape");
  const size_t total = (size_t)B * H * W;
  k_cnn_output<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_net, U, flags, scale, p_out, U_out,
                                                                                 B, H, W, apply_wall_bcs);
  fnx_count_launches(1);
  FNX_CUDA_TRY("fluidnet_output", cudaGetLastError());
  return FNX_OK;
}

// the two stencils around the network of the slice-wise 3-D projection (lib/model.py FluidNet.forward_fields_3d)
int fnx_fluidnet_input_3d(const float* U, const float* flags, const float* scale, float* x, int B, int D, int H, int W,
                          void* stream) {
  if (B < 1 || D < 2 || H < 2 || W < 2) return fnx_set_error(FNX_ERR_ARG, "fluidnet_input_3d: bad shape");
  const size_t total = (size_t)B * D * H * W;
  k_cnn_input_3d<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(U, flags, scale, x, B, D, H, W);
  fnx_count_launches(1);
  FNX_CUDA_TRY("fluidnet_input_3d", cudaGetLastError());
  return FNX_OK;
}

int fnx_fluidnet_output_3d(const float* p_net, const float* U, const float* flags, const float* scale, float* p_out,
                           float* U_out, int B, int D, int H, int W, void* stream) {
  if (B < 1 || D < 2 || H < 2 || W < 2) return fnx_set_error(FNX_ERR_ARG, "fluidnet_output_3d: bad shape");
  const size_t total = (size_t)B * D * H * W;
  k_cnn_output_3d<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_net, U, flags, scale, p_out, U_out,
                                                                                    B, D, H, W);
  fnx_count_launches(1);
  FNX_CUDA_TRY("fluidnet_output_3d", cudaGetLastError());
  return FNX_OK;
}

}  // extern "C"
