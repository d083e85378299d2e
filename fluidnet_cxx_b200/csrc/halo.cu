// halo.cu -- halo exchange of the slab-decomposed step over NVLink peer memory, as ONE kernel.
//
// The reference has no multi-device path.  Each rank's state arrays live in
// symmetric memory that every peer maps (torch.distributed symmetric memory provides allocation and
// rendezvous; see lib/slab.py).  An exchange is a single launch on every rank:
//   1. all CTAs copy this rank's boundary rows STRAIGHT INTO the neighbours' ghost rows (16-byte
//      stores through the peer mapping: NVLink / NVSwitch writes), then fence at system scope;
//   2. the last CTA to finish raises this exchange site's flag in every neighbour's memory
//      (st.release.sys) and then waits until every neighbour has raised the flag in OURS
//      (ld.acquire.sys).
// When the kernel completes, the ghost rows hold the neighbours' data and the neighbours have been told
// about ours: no pack / unpack kernels, no host call, no NCCL in the data path, so the whole time step
// (stencil kernels + exchanges) is one CUDA graph.  Flags are per exchange site and count upwards (an
// epoch kept on the device), so graph replays need no reset.  Write-after-read safety follows from the
// schedule (lib/slab.py): a rank only ever pushes into rows its neighbour does not touch between the
// neighbour's previous signal to it and the neighbour's wait for this push.
//
// Every wait is bounded: a flag that does not arrive within FNX_HALO_TIMEOUT_CLOCKS traps, turning a
// lost peer into a CUDA error on this rank instead of a hang.
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"
#include "host_util.h"

namespace fnx {

#ifndef FNX_HALO_TIMEOUT_CLOCKS
#define FNX_HALO_TIMEOUT_CLOCKS 20000000000LL   // ~10 s at 1.9 GHz: a step lasts milliseconds
#endif

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_halo_exchange(const fnx_halo_desc d) {
  // ---- 1. push ----
  for (int k = 0; k < d.n_peers; k++) {
    const size_t n4 = d.count[k] >> 2;   // float4 units (counts are multiples of 4, pointers 16-byte aligned)
    for (int f = 0; f < d.n_fields; f++) {
      const float4* __restrict__ src = reinterpret_cast<const float4*>(d.src[k][f]);
      float4* __restrict__ dst = reinterpret_cast<float4*>(d.dst[k][f]);
      for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x)
        dst[e] = src[e];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  // ---- 2. signal + wait (last CTA only) ----
  const unsigned prev = atomicAdd(d.done, 1u);
  if (prev != gridDim.x - 1) return;
  __threadfence_system();
  const unsigned e = *d.epoch + 1u;
  for (int k = 0; k < d.n_peers; k++) st_release_sys(d.flag_out[k], e);
  const long long t0 = clock64();
  for (int k = 0; k < d.n_peers; k++) {
    unsigned spins = 0;
    // (int) difference: correct across the 2^32 wrap of the epoch
    while ((int)(ld_acquire_sys(d.flag_in[k]) - e) < 0) {
      if ((++spins & 255u) == 0 && clock64() - t0 > FNX_HALO_TIMEOUT_CLOCKS) __trap();
    }
  }
  *d.epoch = e;
  *d.done = 0u;
  __threadfence();
}

}  // namespace fnx

extern "C" int fnx_halo_exchange(const fnx_halo_desc* d, void* stream) {
  if (!d || d->n_peers < 0 || d->n_peers > FNX_HALO_MAX_PEERS || d->n_fields < 0 || d->n_fields > FNX_HALO_MAX_FIELDS)
    return fnx_set_error(FNX_ERR_ARG, "halo_exchange: bad descriptor");
  if (!d->epoch || !d->done) return fnx_set_error(FNX_ERR_ARG, "halo_exchange: epoch / done counters missing");
  size_t most = 0;
  for (int k = 0; k < d->n_peers; k++) {
    if (d->count[k] % 4) return fnx_set_error(FNX_ERR_ARG, "halo_exchange: counts must be multiples of 4 floats");
    if (!d->flag_out[k] || !d->flag_in[k]) return fnx_set_error(FNX_ERR_ARG, "halo_exchange: flag pointers missing");
    for (int f = 0; f < d->n_fields; f++)
      if (((uintptr_t)d->src[k][f] | (uintptr_t)d->dst[k][f]) & 15)
        return fnx_set_error(FNX_ERR_ARG, "halo_exchange: rows must be 16-byte aligned");
    most = d->count[k] * d->n_fields > most ? d->count[k] * d->n_fields : most;
  }
  if (d->n_peers == 0) return FNX_OK;
  // enough CTAs to keep the NVLink stores in flight, few enough that the "last CTA" hand-off is cheap
  size_t want = (most / 4 + 256 * 8 - 1) / (256 * 8);
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 64 ? 64 : want));
  fnx::k_halo_exchange<<<grid, 256, 0, (cudaStream_t)stream>>>(*d);
  fnx_count_launches(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "halo_exchange: %s", cudaGetErrorString(e));
  return FNX_OK;
}
