/*
 * fluidstep.h -- C-ABI of the B200-native per-timestep fluid path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's
 * FFI for this path binds.  In jolibrain/fluidnet_cxx that FFI is the pybind
 * module `fluidnet_cpp` (pytorch/lib/fluid/cpp/fluids_init.cpp:1009-1014) plus the
 * pure-torch stencil functions of `lib.fluid` (pytorch/lib/fluid/__init__.py:1-14).
 * Each function cites the reference interface it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to a contiguous fp32 tensor in the
 *    reference layout (B, C, D, H, W); flags are fp32-encoded Manta cell types
 *    (pytorch/lib/fluid/cell_type.py:5-14);
 *  - `is3d` selects 2 (D must be 1) or 3 velocity channels;
 *  - `stream` is a cudaStream_t (0 = legacy default stream); all work is
 *    enqueued on it and the call returns without synchronising unless stated;
 *  - nothing is allocated inside: outputs and workspaces are caller-provided
 *    (`fnx_*_workspace` returns the byte size a call needs);
 *  - return value 0 = ok, negative = error (see fnx_last_error()).
 *  - in-place ops mutate `U` exactly as the reference's Python ops do.
 *
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef FLUIDSTEP_H_
#define FLUIDSTEP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNX_OK 0
#define FNX_ERR_ARG (-1)   /* bad shape / unsupported argument (reference: AssertionError / AT_ERROR) */
#define FNX_ERR_CUDA (-2)  /* CUDA runtime error, message in fnx_last_error() */
#define FNX_ERR_WORKSPACE (-3)

#define FNX_METHOD_EULER 0      /* 'eulerFluidNet'      advect_type.cpp:5-16 */
#define FNX_METHOD_MACCORMACK 1 /* 'maccormackFluidNet' */

/* library / build information */
const char *fnx_last_error(void);
const char *fnx_build_info(void);  /* "sm_100a ..." */
int fnx_abi_version(void);
/* number of kernels this library has launched so far in this process (monotonic) */
long long fnx_launch_count(void);

/* ---- advection: fluidnet_cpp.advect_scalar / advect_vel ------------------- */
/* fluids_init.h:73-100 / fluids_init.cpp:265-382  (wrapper advection.py:14-66).
 * dst (B,1,D,H,W) is a new tensor; src, U, flags are not modified. */
size_t fnx_advect_scalar_workspace(int B, int D, int H, int W);
int fnx_advect_scalar(float dt, const float *src, const float *U, const float *flags, float *dst,
                      int B, int D, int H, int W, int is3d, int method, int boundary_width,
                      int sample_outside_fluid, float maccormack_strength, void *workspace,
                      size_t workspace_bytes, void *stream);

/* fluids_init.h:102-124 / fluids_init.cpp:656-807 (wrapper advection.py:68-118).
 * dst (B,2|3,D,H,W) is a new tensor; orig may alias U (self-advection). */
size_t fnx_advect_vel_workspace(int B, int D, int H, int W, int is3d);
int fnx_advect_vel(float dt, const float *orig, const float *U, const float *flags, float *dst,
                   int B, int D, int H, int W, int is3d, int method, int boundary_width,
                   float maccormack_strength, void *workspace, size_t workspace_bytes,
                   void *stream);

/* ---- pressure solve: fluidnet_cpp.solve_linear_system ---------------------- */
/* fluids_init.h:126-142 / fluids_init.cpp:809-1004 (wrapper solve_linear_sys.py:4-40).
 * p (B,1,D,H,W) out (p0 = 0 as in the reference); residual: 1 device float out (may be NULL when
 * p_tol <= 0: the residual pass of the last iteration is skipped).
 * p_tol > 0 makes the call poll a device flag every few iterations (the
 * reference syncs every iteration); p_tol <= 0 never synchronises.
 * `iters_run` (host int, may be NULL) receives the iterations executed
 * (only meaningful when p_tol > 0; else max_iter). */
size_t fnx_jacobi_workspace(int B, int D, int H, int W, int max_iter);
int fnx_solve_linear_system_jacobi(const float *flags, const float *div, float *p, float *residual,
                                   int B, int D, int H, int W, int is3d, float p_tol,
                                   int max_iter, int *iters_run, void *workspace,
                                   size_t workspace_bytes, void *stream);

/* `iters` Jacobi iterations of the same system continued from p_init (NULL = start from p = 0):
 * p_init -> p, no residual.  Building block of the slab-decomposed solver (halo exchange of p
 * between chunks of iterations); p_init must not alias p.  [row_begin, row_end) restricts the
 * rows written (0,0 = all): rows outside are read as they are (stale halo), never written. */
int fnx_jacobi_iterate(const float *flags, const float *div, const float *p_init, float *p, int B,
                       int D, int H, int W, int is3d, int iters, int row_begin, int row_end,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Residual-terminated Jacobi across slabs.  The reference evaluates max_b ||p - p_prev||_2 after every
 * iteration and stops below p_tol (pytorch/lib/fluid/cpp/fluids_init.cpp:958-990).  `iters` iterations,
 * one per launch, continued from p_init; ssq[it * B + b] (device, iters*B doubles, zeroed by the call)
 * = sum over rows [own_begin, own_end) of (p_it - p_{it-1})^2: this rank's share of iteration it's squared
 * residual, to be summed over ranks by the caller.  workspace: B*D*H*W floats.  Never synchronises. */
int fnx_jacobi_iterate_resid(const float *flags, const float *div, const float *p_init, float *p, int B,
                             int D, int H, int W, int is3d, int iters, int row_begin, int row_end,
                             int own_begin, int own_end, double *ssq, void *workspace,
                             size_t workspace_bytes, void *stream);

/* 2-D, arrays that hold rows [held_row_begin, held_row_end) of the H x W grid only (a slab): */
int fnx_jacobi_iterate_held(const float *flags, const float *div, const float *p_init, float *p,
                            int B, int H, int W, int iters, int row_begin, int row_end,
                            int held_row_begin, int held_row_end, void *workspace,
                            size_t workspace_bytes, void *stream);

/* the tile masks of one launch shape (flags are static during a simulation), and a launch (1..8 iterations) using them */
size_t fnx_jacobi_tilemask_bytes(int B, int rows, int W);
int fnx_jacobi_tilemask_held(const float *flags, int B, int H, int W, int row_begin, int row_end,
                             int held_row_begin, int held_row_end, void *masks, size_t mask_bytes,
                             void *stream);
int fnx_jacobi_iterate_held_masked(const float *flags, const float *div, const float *p_init, float *p,
                                   int B, int H, int W, int iters, int row_begin, int row_end,
                                   int held_row_begin, int held_row_end, const void *masks,
                                   size_t mask_bytes, void *stream);

/* ---- halo exchange over NVLink peer memory (no reference counterpart: SURVEY.md section 8e) ----
 * One kernel per exchange: copy `count[k]` floats of each of `n_fields` fields from src[k][f] (this
 * GPU) to dst[k][f] (peer-mapped memory of neighbour k), then raise *flag_out[k] (in neighbour k's
 * memory) to the next epoch and wait until every *flag_in[k] (in this GPU's memory, raised by
 * neighbour k's matching call) has reached it.  `epoch` / `done`: two zero-initialised device words
 * owned by this exchange site.  Capturable in CUDA graphs; waits are bounded (trap on timeout). */
#define FNX_HALO_MAX_PEERS 8
#define FNX_HALO_MAX_FIELDS 4
typedef struct fnx_halo_desc {
  int n_peers, n_fields;
  const float *src[FNX_HALO_MAX_PEERS][FNX_HALO_MAX_FIELDS];
  float *dst[FNX_HALO_MAX_PEERS][FNX_HALO_MAX_FIELDS];
  size_t count[FNX_HALO_MAX_PEERS];
  unsigned *flag_out[FNX_HALO_MAX_PEERS];
  const unsigned *flag_in[FNX_HALO_MAX_PEERS];
  unsigned *epoch, *done;
} fnx_halo_desc;
int fnx_halo_exchange(const fnx_halo_desc *d, void *stream);

/* ---- lib.fluid stencils ----------------------------------------------------- */
/* velocity_divergence.py:4-74 : div (B,1,D,H,W) out */
int fnx_velocity_divergence(const float *U, const float *flags, float *div, int B, int D, int H,
                            int W, int is3d, void *stream);
/* velocity_update.py:6-162 : U -= grad(p), in place */
int fnx_velocity_update(const float *pressure, float *U, const float *flags, int B, int D, int H,
                        int W, int is3d, void *stream);
/* set_wall_bcs.py:4-86 : in place */
int fnx_set_wall_bcs(float *U, const float *flags, int B, int D, int H, int W, int is3d,
                     void *stream);
/* source_terms.py:6-116 : in place; gravity = 3 host floats */
int fnx_add_buoyancy(float *U, const float *flags, const float *density, const float *gravity3,
                     float rho_star, float dt, int B, int D, int H, int W, int is3d, void *stream);
/* source_terms.py:122-219 : in place */
int fnx_add_gravity(float *U, const float *flags, const float *gravity3, float dt, int B, int D,
                    int H, int W, int is3d, void *stream);
/* viscosity.py:7-70 addViscosity (2-D): in place on U (B,2,1,H,W); workspace >= B*2*H*W floats.
 * dt and viscosity are the Python floats of the call: their product is formed in double, as there. */
int fnx_add_viscosity(float *U, const float *flags, double dt, double viscosity, int B, int H, int W,
                      void *workspace, size_t workspace_bytes, void *stream);
/* advection.py:9-12 correctScalar: src += dt*0.5*src*div on Fluid cells, in place */
int fnx_correct_scalar(float *src, const float *div, const float *flags, double dt, size_t count,
                       void *stream);
/* flags_to_occupancy.py:6-19 */
int fnx_flags_to_occupancy(const float *flags, float *occupancy, size_t count, void *stream);
/* simulate.py:4-26 setConstVals : x = x*inv_mask + bc, in place */
int fnx_set_const_vals(float *x, const float *inv_mask, const float *bc, size_t count,
                       void *stream);
/* util.py:5-49 emptyDomain : border ring (width bnd) = Obstacle, inside = Fluid */
int fnx_empty_domain(float *flags, int B, int D, int H, int W, int is3d, int bnd, void *stream);
/* grid.py:7-30 getCentered : out (B,3,D,H,W) cell-centred velocity */
int fnx_get_centered(const float *U, float *out, int B, int D, int H, int W, int is3d,
                     void *stream);

/* The drivers' output step (pytorch/plume.py:238-263 and :330-423; rayleighTaylor.py has the same block) in
 * one pass: out (B, FNX_OUTPUT_PLANES, D, H, W) =
 *   0 velocityDivergence(U, flags)          (velocity_divergence.py:4-74)
 *   1,2,3 getCentered(U) x, y, z            (grid.py:7-30)         4 its norm over the components
 *   5,6 centred density gradient x, y       (plume.py:343-350, 2-D; 0 in 3-D)
 *   7,8 centred pressure gradient x, y      (plume.py:352-359)
 *   9 pressure
 * mask_obstacles: planes 1-4 and 9 hold NaN in Obstacle cells (the drivers' masked_array .filled(nan)).
 * density / pressure may be NULL (their planes are 0).  One buffer -> one device-to-host copy. */
#define FNX_OUTPUT_PLANES 10
int fnx_output_fields(const float *U, const float *flags, const float *density, const float *pressure,
                      float *out, int B, int D, int H, int W, int is3d, int mask_obstacles,
                      void *stream);

/* ---- fused fast path behind lib.simulate (simulate.py:28-171) --------------- */
/* The standard inviscid sequence of both reference drivers
 *   advectScalar -> advectVelocity -> setConstVals -> addBuoyancy/addGravity ->
 *   [setWallBcs] -> setConstVals -> velocityDivergence -> (Jacobi | CNN)
 *   -> velocityUpdate -> [setWallBcs] -> setConstVals
 * in four kernels + the pressure solve, with per-cell arithmetic identical to the
 * individual entry points.  density_out / U_out may alias density_in / U_in
 * (in place) or be new tensors; masks may be NULL (no imposed values). */
typedef struct fnx_step_params {
  float dt, maccormack_strength;
  int sample_outside_fluid;
  int use_buoyancy, use_gravity;
  float buoyancy3[3]; /* gravityVec * (-buoyancyScale), simulate.py:101-105 */
  float gravity3[3];  /* gravityVec * (-gravityScale),  simulate.py:109-114 */
  float rho_star;
  int jacobi_iters;          /* fnx_step_jacobi */
  int apply_wall_bcs;        /* simulate.py:120-123: 1 in the jacobi branch, 0 ahead of the CNN */
  int density_const_passes;  /* further setConstVals passes folded into density_out (simulate.py:133,168) */
  /* slab window of the domain-decomposed step: only rows [row_begin, row_end) of the flattened
   * (D*H) row space are computed (arrays stay global-sized, coordinates stay global, so the rows
   * computed are bit-identical to the single-GPU step); 0,0 = the whole grid */
  int row_begin, row_end;
  /* 2-D only: the array arguments hold rows [held_row_begin, held_row_end) of the global H x W grid
   * (pointers address the first held row; a 2-channel field holds its channels held-rows apart; the
   * workspace and mask_rows are sized for the held rows).  The window must lie >= 3 rows inside the held
   * rows on every side that is not a grid edge.  0,0 = the arrays hold the whole grid. */
  int held_row_begin, held_row_end;
} fnx_step_params;

size_t fnx_step_workspace(int B, int D, int H, int W, int is3d);
/* rows[b*D*H + row]: bit0 = the U masks of that grid row differ from (InvMask=1, BC=0), bit1 =
 * the density masks do.  Lets the step kernels skip mask loads on identity rows (the plume
 * inlet touches 4 rows).  Recompute whenever a mask tensor changes. */
int fnx_mask_rows(const float *UBC, const float *UBCInvMask, const float *densityBC,
                  const float *densityBCInvMask, unsigned char *rows, int B, int D, int H, int W,
                  int is3d, void *stream);
/* advection + BCs + forces + [wall BCs] + BCs + divergence.  div may be NULL. */
int fnx_step_advect_forces_div(const fnx_step_params *prm, const float *density_in,
                               const float *U_in, const float *flags, const float *UBC,
                               const float *UBCInvMask, const float *densityBC,
                               const float *densityBCInvMask, const unsigned char *mask_rows,
                               float *density_out, float *U_out, float *div, int B, int D, int H,
                               int W, int is3d, void *workspace, size_t workspace_bytes,
                               void *stream);
/* velocityUpdate + [setWallBcs] + setConstVals, in place on U */
int fnx_step_project_bcs(const float *pressure, float *U, const float *flags, const float *UBC,
                         const float *UBCInvMask, const unsigned char *mask_rows,
                         int apply_wall_bcs, int B, int D, int H, int W, int is3d, void *stream);
int fnx_step_project_bcs_rows(const float *pressure, float *U, const float *flags, const float *UBC,
                              const float *UBCInvMask, const unsigned char *mask_rows,
                              int apply_wall_bcs, int B, int D, int H, int W, int is3d,
                              int row_begin, int row_end, void *stream);
/* the same on arrays that hold rows [held_row_begin, held_row_end) only (2-D; see fnx_step_params) */
int fnx_step_project_bcs_held(const float *pressure, float *U, const float *flags, const float *UBC,
                              const float *UBCInvMask, const unsigned char *mask_rows,
                              int apply_wall_bcs, int B, int D, int H, int W, int is3d,
                              int row_begin, int row_end, int held_row_begin, int held_row_end,
                              void *stream);
/* whole Jacobi step; p and residual as in fnx_solve_linear_system_jacobi (p_tol = 0) */
int fnx_step_jacobi(const fnx_step_params *prm, const float *density_in, const float *U_in,
                    const float *flags, const float *UBC, const float *UBCInvMask,
                    const float *densityBC, const float *densityBCInvMask,
                    const unsigned char *mask_rows, float *density_out, float *U_out, float *p,
                    float *residual, int B, int D, int H, int W, int is3d, void *workspace,
                    size_t workspace_bytes, void *stream);

/* ---- FluidNet / MultiScaleNet forward (model.py:76-227, multi_scale_net.py:101-127) */
/* unbiased std over all elements of x per batch row, clamped below by `threshold`
 * (model.py:8-23 _ScaleNet).  scale: B device floats out. */
size_t fnx_scale_std_workspace(int B);
int fnx_scale_std(const float *x, size_t count_per_batch, int B, float threshold, float *scale,
                  void *workspace, size_t workspace_bytes, void *stream);
/* nn.Conv2d(Cin, Cout, ksize, padding=ksize/2), NCHW fp32, optional fused ReLU
 * (multi_scale_net.py:31-39,55-67,83-95,116).  The result is written into channels
 * [y_channel_offset, +Cout) of a (N, y_channels_total, H, W) tensor. */
int fnx_conv2d(const float *x, const float *weight, const float *bias, float *y, int N, int Cin,
               int H, int W, int Cout, int ksize, int relu, int y_channels_total,
               int y_channel_offset, void *stream);
/* F.upsample(x, (Ho, Wo), mode='bilinear') with align_corners=False (multi_scale_net.py:121-125),
 * written into a channel window of y like fnx_conv2d (fuses the torch.cat) */
int fnx_resize_bilinear(const float *x, float *y, int N, int C, int H, int W, int Ho, int Wo,
                        int y_channels_total, int y_channel_offset, void *stream);
/* ---- tensor-core path of the MultiScaleNet convolutions (multi_scale_net.py:101-127) ----------
 * nn.Conv2d(Cin, Cout, k, padding=k/2) for k = 3 (Cin, Cout <= 128) and k = 5 (Cout <= 32) as a
 * tcgen05 implicit GEMM with fp32-equivalent accuracy (two-term fp16 expansion, three MMAs per K
 * step, fp32 TMEM accumulation; channels zero-padded to MMA granularity inside).  Activations travel between these layers in the "split chunked"
 * layout: two fp16 planes (hi, lo) [C/8][H+2P][W+2P][8] with a zero border of P pixels that the
 * caller zeroes ONCE (the kernels never write it) plus a device-side fnx_act_meta. */
#define FNX_TC_PAD 2
typedef struct fnx_act_meta {
  unsigned amax_bits; /* fp32 bit pattern of max|a| over the tensor (atomicMax target; zero it first) */
  float scale;        /* power of two the stored expansion is multiplied by */
} fnx_act_meta;
size_t fnx_tc_act_bytes(int C, int H, int W);
size_t fnx_tc_weight_bytes(int Cin, int Cout, int ksize);
/* weight (Cout, Cin, k, k) fp32 -> packed split-fp16 slots; w_scale = power of two with max|w|*w_scale <= 2^14 */
int fnx_tc_pack_weights(const float *weight, int Cin, int Cout, int ksize, float w_scale, void *out,
                        void *stream);
/* tuning aid: while buf != NULL every fnx_conv_tc launch writes, per CTA b, buf[4b..4b+3] = clocks its MMA warp
 * spent waiting for {activation stages, weight slots, accumulators handed back by the epilogue} and its total */
int fnx_tc_set_debug(long long *buf);
/* meta->amax_bits = max(meta->amax_bits, max|x|) */
int fnx_tc_amax(const float *x, size_t n, fnx_act_meta *meta, void *stream);
/* fp32 NCHW (one image) -> split chunked, scale chosen from in_meta->amax_bits */
int fnx_tc_pack_split(const float *x, int C, int H, int W, const fnx_act_meta *in_meta, void *y,
                      fnx_act_meta *out_meta, void *stream);
int fnx_tc_unpack_split(const void *x, const fnx_act_meta *meta, int C, int H, int W, float *y, void *stream);
/* out_mode 0: y split chunked (out_meta required, amax_bits pre-zeroed, Cout % 16 == 0); out_mode 1:
 * y fp32 NCHW channel window like fnx_conv2d.  w_norm = max_n sum|w[n]|, b_max = max|bias| (the
 * fp16 range bound of the output). */
int fnx_conv_tc(const void *x, const fnx_act_meta *in_meta, const void *w_packed, const float *bias,
                int Cin, int Cout, int ksize, int H, int W, int relu, float w_scale, float w_norm,
                float b_max, int out_mode, void *y, fnx_act_meta *out_meta, int y_channels_total,
                int y_channel_offset, void *stream);

/* MultiScaleNet.forward in one call (multi_scale_net.py:101-127): x (N, data_channels, H, W) ->
 * y (N, 1, H, W).  Layers with w_tc != NULL and a supported shape run on the tensor cores, the
 * others on the fp32 direct kernel.  The workspace must have been zeroed once with
 * fnx_msnet_workspace_init for this (H, W) and not written by anyone else since. */
typedef struct fnx_conv_layer {
  const float *weight; /* (Cout, Cin, k, k) fp32 */
  const float *bias;   /* (Cout) fp32 */
  const void *w_tc;    /* fnx_tc_pack_weights output, or NULL */
  int cin, cout, ksize, relu;
  float w_scale, w_norm, b_max;
  int w_replicas;      /* w_tc holds this many identical copies, fnx_tc_weight_bytes apart (0/1 = one) */
} fnx_conv_layer;
typedef struct fnx_msnet_plan {
  int data_channels;
  fnx_conv_layer quarter[4], half[6], full[6], final_conv;
} fnx_msnet_plan;
size_t fnx_msnet_workspace(const fnx_msnet_plan *plan, int H, int W);
/* workspace of a forward over N images: with the all-tensor-core plan the batch runs as ONE launch per layer (the
 * images stacked into a tall image, each keeping its own zero padding; activation scales shared by the batch) */
size_t fnx_msnet_workspace_n(const fnx_msnet_plan *plan, int N, int H, int W);
int fnx_msnet_workspace_init(void *workspace, size_t workspace_bytes, void *stream);
int fnx_msnet_forward(const fnx_msnet_plan *plan, const float *x, float *y, int N, int H, int W,
                      void *workspace, size_t workspace_bytes, void *stream);

/* optional per-layer timing of fnx_msnet_forward (bench.py's roofline): while enabled, every conv
 * launch is bracketed by CUDA events on its stream; fetch synchronises them, returns the number of
 * records since the last fetch / enable and clears them. */
typedef struct fnx_profile_rec {
  int cin, cout, ksize, h, w, tensor; /* tensor: 1 = tcgen05 kernel, 0 = fp32 direct kernel */
  float ms;
} fnx_profile_rec;
int fnx_profile_enable(int enable);
int fnx_profile_fetch(fnx_profile_rec *out, int capacity);

/* net input of the shipped ScaleNet: x = [velocityDivergence(U, flags)/scale, flagsToOccupancy(flags)]
 * (*_saved.py:135-177); U (B,2,1,H,W), x (B,2,H,W) */
int fnx_fluidnet_input(const float *U, const float *flags, const float *scale, float *x, int B,
                       int H, int W, void *stream);
/* post-processing of the wrapper (*_saved.py:221-232): U/scale -> velocityUpdate(p_net) ->
 * *scale -> [setWallBcs] ; p_out = p_net*scale.  apply_wall_bcs = 0 gives the field the periodic
 * seam copy of *_saved.py:228-237 reads (the velocity before setWallBcs). */
int fnx_fluidnet_output(const float *p_net, const float *U, const float *flags, const float *scale,
                        float *p_out, float *U_out, int B, int H, int W, int apply_wall_bcs,
                        void *stream);

/* The same two stencils for the SLICE-WISE 3-D projection this package defines for BASELINE configs[4] (no reference
 * counterpart: the reference model is 2-D only, pytorch/lib/model.py:93).  input: x ((B*D), 2, H, W) = one image
 * per z-slice: [velocityDivergence_3D(U, flags) / scale[b], flagsToOccupancy(flags)].  output: p_net ((B*D),1,H,W)
 * -> U_out = setWallBcs_3D((U/s with the IN-PLANE pressure gradient subtracted from Ux, Uy; Uz kept) * s),
 * p_out = p_net * s. */
int fnx_fluidnet_input_3d(const float *U, const float *flags, const float *scale, float *x, int B, int D,
                          int H, int W, void *stream);
int fnx_fluidnet_output_3d(const float *p_net, const float *U, const float *flags, const float *scale,
                           float *p_out, float *U_out, int B, int D, int H, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDSTEP_H_ */
