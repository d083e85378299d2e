// step.cu -- fused entry points behind lib.simulate (pytorch/lib/simulate.py:28-171) and the
// global reduction of the FluidNet wrapper (_ScaleNet, pytorch/lib/model.py:8-23).
// Compiled with -fmad=false.
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"
#include "advect_device.cuh"
#include "fluid_common.cuh"
#include "host_util.h"
#include "stencil_device.cuh"

namespace fnx {

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- unbiased std (model.py:18) ----------------------------------------------------------
// pass 1: per-batch sum -> mean; pass 2: sum (x-mean)^2; both in double (torch accumulates the
// fp32 CPU std in double as well), one atomic per block.
__global__ void __launch_bounds__(256)
    k_sum(const float* __restrict__ x, size_t count, double* __restrict__ out, const double* __restrict__ sum_in,
          int pass) {
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * count;
  double mean = 0.0;
  if (pass == 1) mean = sum_in[b] / (double)count;
  double acc = 0.0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)__ldg(xb + q);
    if (pass == 0) acc += v;
    else { const double d = v - mean; acc += d * d; }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ double wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += wsum[w];
    atomicAdd(out + b, t);
  }
}

__global__ void k_std_finish(const double* __restrict__ ssq, size_t count, int B, float threshold,
                             float* __restrict__ scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = (float)sqrt(ssq[b] / (double)(count - 1));
  scale[b] = s < threshold ? threshold : s;  // torch.clamp(std, min=threshold); NaN stays NaN
}

}  // namespace fnx

using namespace fnx;

#define FNX_TRY(call)            \
  do {                           \
    int e_ = (call);             \
    if (e_ != FNX_OK) return e_; \
  } while (0)

#define FNX_CUDA_TRY(who, call)                                                      \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); \
  } while (0)

extern "C" {

size_t fnx_scale_std_workspace(int B) { return align256(2 * (size_t)B * sizeof(double)); }

int fnx_scale_std(const float* x, size_t count_per_batch, int B, float threshold, float* scale, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (B < 1 || count_per_batch < 2) return fnx_set_error(FNX_ERR_ARG, "scale_std: need B>=1 and >=2 elements");
  if (!workspace || workspace_bytes < fnx_scale_std_workspace(B))
    return fnx_set_error(FNX_ERR_WORKSPACE, "scale_std: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* sum = (double*)workspace;
  double* ssq = sum + B;
  FNX_CUDA_TRY("scale_std", cudaMemsetAsync(sum, 0, 2 * (size_t)B * sizeof(double), st));
  size_t want = (count_per_batch + 256 * 8 - 1) / (256 * 8);
  unsigned nblk = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  dim3 grid(nblk, B);
  k_sum<<<grid, 256, 0, st>>>(x, count_per_batch, sum, nullptr, 0);
  k_sum<<<grid, 256, 0, st>>>(x, count_per_batch, ssq, sum, 1);
  k_std_finish<<<(B + 63) / 64, 64, 0, st>>>(ssq, count_per_batch, B, threshold, scale);
  fnx_count_launches(3);
  FNX_CUDA_TRY("scale_std", cudaGetLastError());
  return FNX_OK;
}

// ---- fused step (v1: the standard sequence, one entry point) --------------------------------
size_t fnx_step_workspace(int B, int D, int H, int W, int is3d) {
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  size_t s = 0;
  s += align256(n * sizeof(float));        // advected density
  s += align256(n * nc * sizeof(float));   // advected velocity
  s += align256(fnx_advect_scalar_workspace(B, D, H, W));
  s += align256(fnx_advect_vel_workspace(B, D, H, W, is3d));
  s += align256(n * sizeof(float));        // div (fnx_step_jacobi)
  s += align256(fnx_jacobi_workspace(B, D, H, W, 1));
  return s;
}

int fnx_step_advect_forces_div(const fnx_step_params* prm, float* density, float* U, const float* flags,
                               const float* UBC, const float* UBCInvMask, const float* densityBC,
                               const float* densityBCInvMask, float* div, int B, int D, int H, int W, int is3d,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!prm) return fnx_set_error(FNX_ERR_ARG, "step: null params");
  if (!workspace || workspace_bytes < fnx_step_workspace(B, D, H, W, is3d))
    return fnx_set_error(FNX_ERR_WORKSPACE, "step: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  char* ws = (char*)workspace;
  float* rho_adv = (float*)ws; ws += align256(n * sizeof(float));
  float* U_adv = (float*)ws; ws += align256(n * nc * sizeof(float));
  void* ws_s = ws; const size_t ws_s_bytes = align256(fnx_advect_scalar_workspace(B, D, H, W)); ws += ws_s_bytes;
  void* ws_v = ws; const size_t ws_v_bytes = align256(fnx_advect_vel_workspace(B, D, H, W, is3d));

  // simulate.py:75-94: both advections read the velocity field of the previous step
  FNX_TRY(fnx_advect_scalar(prm->dt, density, U, flags, rho_adv, B, D, H, W, is3d, FNX_METHOD_MACCORMACK, 1,
                            prm->sample_outside_fluid, prm->maccormack_strength, ws_s, ws_s_bytes, stream));
  FNX_TRY(fnx_advect_vel(prm->dt, U, U, flags, U_adv, B, D, H, W, is3d, FNX_METHOD_MACCORMACK, 1,
                         prm->maccormack_strength, ws_v, ws_v_bytes, stream));
  FNX_CUDA_TRY("step", cudaMemcpyAsync(density, rho_adv, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  FNX_CUDA_TRY("step", cudaMemcpyAsync(U, U_adv, n * nc * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const bool ubc = UBC && UBCInvMask, rbc = densityBC && densityBCInvMask;
  // simulate.py:96 setConstVals
  if (ubc) FNX_TRY(fnx_set_const_vals(U, UBCInvMask, UBC, n * nc, stream));
  if (rbc) FNX_TRY(fnx_set_const_vals(density, densityBCInvMask, densityBC, n, stream));
  // simulate.py:98-115 forces
  if (prm->use_buoyancy)
    FNX_TRY(fnx_add_buoyancy(U, flags, density, prm->buoyancy3, prm->rho_star, prm->dt, B, D, H, W, is3d, stream));
  if (prm->use_gravity) FNX_TRY(fnx_add_gravity(U, flags, prm->gravity3, prm->dt, B, D, H, W, is3d, stream));
  // simulate.py:120-133 wall BCs + setConstVals
  FNX_TRY(fnx_set_wall_bcs(U, flags, B, D, H, W, is3d, stream));
  if (ubc) FNX_TRY(fnx_set_const_vals(U, UBCInvMask, UBC, n * nc, stream));
  if (rbc) FNX_TRY(fnx_set_const_vals(density, densityBCInvMask, densityBC, n, stream));
  // simulate.py:145
  if (div) FNX_TRY(fnx_velocity_divergence(U, flags, div, B, D, H, W, is3d, stream));
  return FNX_OK;
}

int fnx_step_project_bcs(const float* pressure, float* U, const float* flags, const float* UBC,
                         const float* UBCInvMask, int B, int D, int H, int W, int is3d, void* stream) {
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  FNX_TRY(fnx_velocity_update(pressure, U, flags, B, D, H, W, is3d, stream));  // simulate.py:154
  FNX_TRY(fnx_set_wall_bcs(U, flags, B, D, H, W, is3d, stream));               // :159
  if (UBC && UBCInvMask) FNX_TRY(fnx_set_const_vals(U, UBCInvMask, UBC, n * nc, stream));  // :168
  return FNX_OK;
}

int fnx_step_jacobi(const fnx_step_params* prm, float* density, float* U, const float* flags, float* p,
                    float* residual, const float* UBC, const float* UBCInvMask, const float* densityBC,
                    const float* densityBCInvMask, int B, int D, int H, int W, int is3d, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (!prm) return fnx_set_error(FNX_ERR_ARG, "step: null params");
  if (!workspace || workspace_bytes < fnx_step_workspace(B, D, H, W, is3d))
    return fnx_set_error(FNX_ERR_WORKSPACE, "step: workspace too small");
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  char* ws = (char*)workspace;
  ws += align256(n * sizeof(float)) + align256(n * nc * sizeof(float)) +
        align256(fnx_advect_scalar_workspace(B, D, H, W)) + align256(fnx_advect_vel_workspace(B, D, H, W, is3d));
  float* div = (float*)ws; ws += align256(n * sizeof(float));
  void* ws_j = ws; const size_t ws_j_bytes = align256(fnx_jacobi_workspace(B, D, H, W, 1));
  FNX_TRY(fnx_step_advect_forces_div(prm, density, U, flags, UBC, UBCInvMask, densityBC, densityBCInvMask, div, B,
                                     D, H, W, is3d, workspace, workspace_bytes, stream));
  FNX_TRY(fnx_solve_linear_system_jacobi(flags, div, p, residual, B, D, H, W, is3d, 0.f, prm->jacobi_iters, nullptr,
                                         ws_j, ws_j_bytes, stream));
  FNX_TRY(fnx_step_project_bcs(p, U, flags, UBC, UBCInvMask, B, D, H, W, is3d, stream));
  // simulate.py:168: the last setConstVals also re-applies the density BC
  if (densityBC && densityBCInvMask) FNX_TRY(fnx_set_const_vals(density, densityBCInvMask, densityBC, n, stream));
  return FNX_OK;
}

}  // extern "C"
