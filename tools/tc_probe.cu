// tc_probe.cu -- hardware probe for the tcgen05 descriptor conventions conv_tc.cu relies on:
// K-major no-swizzle operands with an affine row layout (SBO = 128 B, so row m sits at +16*m and a
// shifted convolution tap is just a start-address offset), bulk-copy staging, TMEM lane mapping.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../fluidnet_cxx_b200/csrc/tc_ptx.cuh"

using namespace fnx::tc;

struct Exp {
  const char* name;
  uint32_t a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, N, ksteps, a_kadv, b_kadv, a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128) k_probe(const uint8_t* imgA, const uint8_t* imgB, float* D, Exp e) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((e.a_bytes + 1023) & ~1023u);
  uint32_t ncols = 32;
  while (ncols < e.N) ncols <<= 1;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), e.a_bytes + e.b_bytes);
    bulk_g2s(smem_u32(sA), imgA, e.a_bytes, smem_u32(&bars[0]));
    bulk_g2s(smem_u32(sB), imgB, e.b_bytes, smem_u32(&bars[0]));
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t idesc = idesc_f16_f32acc(128, e.N);
    for (uint32_t s = 0; s < e.ksteps; s++) {
      const uint64_t da = smem_desc_kmajor_noswz(smem_u32(sA) + e.a_off + s * e.a_kadv, e.a_lbo, e.a_sbo);
      const uint64_t db = smem_desc_kmajor_noswz(smem_u32(sB) + e.b_off + s * e.b_kadv, e.b_lbo, e.b_sbo);
      mma_f16_ss(taddr, da, db, idesc, s > 0);
    }
    mma_commit(smem_u32(&bars[1]));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  for (uint32_t c = 0; c < e.N; c += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; i++) D[(size_t)(warp * 32 + lane) * e.N + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, ncols);
}

// ---- throughput probe: `reps` back-to-back MMAs of one shape from no-swizzle K-major operands ----
// distinct A tiles per MMA (start address advanced like a convolution tap) so nothing is collector-cached
__global__ void __launch_bounds__(128) k_mma_rate(int N, int reps, int a_step_bytes, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base;
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint32_t idesc = idesc_f16_f32acc(128, N);
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
      t0 = clock64();
      for (int r = 0; r < reps; r++) {
        const uint64_t da = smem_desc_kmajor_noswz(a0 + (uint32_t)((r & 31) * a_step_bytes), 6 * 2080, 128);
        const uint64_t db = smem_desc_kmajor_noswz(b0, (uint32_t)N * 16, 128);
        mma_f16_ss(taddr + (uint32_t)((r & 1) * 256), da, db, idesc, 1);
      }
      mma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, 512);
}

static float h2f(const uint8_t* img, size_t byte_off) {
  __half h;
  memcpy(&h, img + byte_off, 2);
  return __half2float(h);
}

int main(int argc, char** argv) {
  const uint32_t RP = 130, ROWS = 3;
  const uint32_t A_LBO = ROWS * RP * 16;
  std::vector<Exp> exps = {
      {"base N=64", 0, A_LBO, 128, 0, 64 * 16, 128, 64, 1, 0, 0, 0, 0},
      {"shifted start a+48 b+32", 48, A_LBO, 128, 32, 72 * 16, 128, 64, 1, 0, 0, 0, 0},
      {"row+tap shift (RP+1)*16", (RP + 1) * 16, A_LBO, 128, 0, 64 * 16, 128, 64, 1, 0, 0, 0, 0},
      {"4 k-steps N=128", 16, A_LBO, 128, 0, 128 * 16, 128, 128, 4, 2 * A_LBO, 2 * 128 * 16, 0, 0},
      {"N=32", 0, A_LBO, 128, 0, 32 * 16, 128, 32, 1, 0, 0, 0, 0},
      {"N=16", 0, A_LBO, 128, 0, 16 * 16, 128, 16, 1, 0, 0, 0, 0},
      {"N=256", 0, A_LBO, 128, 0, 256 * 16, 128, 256, 1, 0, 0, 0, 0},
      {"a_sbo=160 (8-wide 2D tile)", 0, 24 * 160, 160, 0, 64 * 16, 128, 64, 1, 0, 0, 0, 0},
      {"a_lbo=16 (chunk1 = next pixel)", 0, 16, 128, 0, 64 * 16, 128, 64, 1, 0, 0, 0, 0},
  };
  if (argc > 1 && !strcmp(argv[1], "rate")) {
    long long* dc;
    cudaMalloc(&dc, 8);
    cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int shapes[] = {16, 32, 64, 128, 256};
    for (int N : shapes)
      for (int step : {0, 16, 2080}) {
        long long hc = 0;
        const int reps = 2048;
        k_mma_rate<<<1, 128, 200 * 1024>>>(N, reps, step, dc);
        k_mma_rate<<<1, 128, 200 * 1024>>>(N, reps, step, dc);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(err)); return 2; }
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d K=16 no-swizzle, A start step %4d B: %7.1f clk/MMA  (tensor floor %d, operand bytes %d -> %d clk at 128 B/clk)\n",
               N, step, (double)hc / reps, N / 2, 4096 + N * 32, (4096 + N * 32) / 128);
      }
    return 0;
  }
  int only = argc > 1 ? atoi(argv[1]) : -1;
  srand(1);
  int fails = 0;
  for (size_t ei = 0; ei < exps.size(); ei++) {
    if (only >= 0 && (int)ei != only) continue;
    Exp e = exps[ei];
    // image sizes: enough for every address the descriptors can reach
    e.a_bytes = 8 * A_LBO + 4096;
    e.b_bytes = e.b_off + e.ksteps * (e.b_kadv ? e.b_kadv : 0) + 2 * e.b_lbo + (e.N / 8) * e.b_sbo + 4096;
    e.a_bytes = (e.a_bytes + 15) & ~15u;
    e.b_bytes = (e.b_bytes + 15) & ~15u;
    std::vector<uint8_t> ha(e.a_bytes), hb(e.b_bytes);
    auto fill = [](std::vector<uint8_t>& v) {
      for (size_t i = 0; i < v.size() / 2; i++) {
        __half h = __float2half((float)((rand() % 2001) - 1000) / 500.0f);
        memcpy(&v[2 * i], &h, 2);
      }
    };
    fill(ha);
    fill(hb);
    uint8_t *da, *db;
    float* dD;
    cudaMalloc(&da, e.a_bytes);
    cudaMalloc(&db, e.b_bytes);
    cudaMalloc(&dD, 128 * e.N * 4);
    cudaMemcpy(da, ha.data(), e.a_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), e.b_bytes, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * e.N * 4);
    size_t smem = ((e.a_bytes + 1023) & ~1023u) + e.b_bytes + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(da, db, dD, e);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("[%zu] %-34s CUDA ERROR %s\n", ei, e.name, cudaGetErrorString(err));
      return 2;
    }
    std::vector<float> hD(128 * e.N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    // model 0: LBO = K-chunk stride, SBO = 8-row group stride; model 1: swapped
    double errs[2] = {0, 0};
    for (int model = 0; model < 2; model++) {
      uint32_t alb = model ? e.a_sbo : e.a_lbo, asb = model ? e.a_lbo : e.a_sbo;
      uint32_t blb = model ? e.b_sbo : e.b_lbo, bsb = model ? e.b_lbo : e.b_sbo;
      for (uint32_t m = 0; m < 128; m++)
        for (uint32_t n = 0; n < e.N; n++) {
          double acc = 0;
          for (uint32_t s = 0; s < e.ksteps; s++)
            for (uint32_t k = 0; k < 16; k++) {
              size_t ao = e.a_off + s * e.a_kadv + (k / 8) * alb + (m / 8) * asb + (m % 8) * 16 + (k % 8) * 2;
              size_t bo = e.b_off + s * e.b_kadv + (k / 8) * blb + (n / 8) * bsb + (n % 8) * 16 + (k % 8) * 2;
              if (ao + 2 > ha.size() || bo + 2 > hb.size()) { acc = 1e30; break; }
              acc += (double)h2f(ha.data(), ao) * (double)h2f(hb.data(), bo);
            }
          double d = fabs(acc - (double)hD[(size_t)m * e.N + n]);
          if (!(d <= errs[model])) errs[model] = d;
        }
    }
    bool ok = errs[0] < 1e-3;
    if (!ok) fails++;
    printf("[%zu] %-34s max|err| model(LBO=K,SBO=MN)=%.3e  swapped=%.3e  D[0][0..3]=%g %g %g %g  %s\n", ei, e.name,
           errs[0], errs[1], hD[0], hD[1], hD[2], hD[3], ok ? "OK" : "MISMATCH");
    cudaFree(da); cudaFree(db); cudaFree(dD);
  }
  printf("probe done, %d mismatches\n", fails);
  return fails ? 1 : 0;
}
