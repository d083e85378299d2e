// step2d.h -- internal interface of the window-aware 2-D step kernels (step2d.cu), used by the
// C-ABI entry points in step.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"

struct fnx_step2d_win {
  int H, W;        // global grid
  int row0, row1;  // rows computed by the launch
  int ya0, ya1;    // rows the arrays hold (ya0 = 0, ya1 = H: whole grid)
};

struct fnx_step2d_masks {  // pointers to the first row held (NULL = no imposed values)
  const float* UBC;
  const float* UBCInv;
  const float* rBC;
  const float* rBCInv;
  const unsigned char* rows;  // per held row (fnx_mask_rows)
};

bool fnx_step2d_supported(int H, int W);
int fnx_step2d_check_window(const fnx_step2d_win& w);
// MacCormack advection of density and velocity + setConstVals.  rho_out gets 1 + rho_passes
// setConstVals passes; rho_mid (may be NULL) receives the one-pass density on rows whose density mask
// is not the identity (the density addBuoyancy reads).
int fnx_step2d_advect(const fnx_step2d_win& w, float dt, float maccormack_strength, int sample_outside, const float* rho,
                      const float* U, const float* flags, const fnx_step2d_masks& mk, int rho_passes, float* rho_out,
                      float* rho_mid, float* U_out, int B, int* tile_ws, cudaStream_t st);
// ints of scratch fnx_step2d_advect's interior fast path needs (tile_ws; NULL = generic kernel only)
size_t fnx_step2d_tile_ws_ints(const fnx_step2d_win& w, int B);
int fnx_step2d_forces_div(const fnx_step2d_win& w, const fnx_step_params* prm, const float* rho, const float* rho_mid,
                          const float* U, const float* flags, const fnx_step2d_masks& mk, float* U_out, float* div, int B,
                          cudaStream_t st);
int fnx_step2d_project(const fnx_step2d_win& w, const float* p, float* U, const float* flags, const fnx_step2d_masks& mk,
                       int wall_bcs, int B, cudaStream_t st);
