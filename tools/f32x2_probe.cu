// f32x2_probe.cu -- does sm_100a's packed fp32 arithmetic (add.rn.f32x2 / mul.rn.f32x2 -> SASS FADD2 / FMUL2) raise the
// fp32 rate per issue slot, and does it round like the scalar instructions?  (Decides whether the blocked Jacobi
// sweep -- 4 FADD + 1 FMUL per cell, issue-bound -- should pair its cells.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_probe tools/f32x2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

constexpr int CH = 16;  // independent chains per thread (pairs in the packed kernel)

__global__ void k_scalar(float* out, float seed, int iters) {
  float a[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; i++) a[i] = seed + i + threadIdx.x;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int i = 0; i < 2 * CH; i++) { a[i] = a[i] + seed; a[i] = a[i] * 0.999f; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * CH; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float seed, int iters) {
  float2 a[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = make_float2(seed + 2 * i + threadIdx.x, seed + 2 * i + 1 + threadIdx.x);
  const float2 sd = make_float2(seed, seed), m = make_float2(0.999f, 0.999f);
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = add2(a[i], sd); a[i] = mul2(a[i], m); }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; i++) { s += a[i].x; s += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// rounding check on random bit patterns
__global__ void k_check(const float* x, const float* y, unsigned* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i + 1 >= n) return;
  float2 a = make_float2(x[2 * i], x[2 * i + 1]), b = make_float2(y[2 * i], y[2 * i + 1]);
  float2 s = add2(a, b), p = mul2(a, b);
  float s0 = a.x + b.x, s1 = a.y + b.y, p0 = a.x * b.x, p1 = a.y * b.y;
  if (__float_as_uint(s.x) != __float_as_uint(s0) || __float_as_uint(s.y) != __float_as_uint(s1) ||
      __float_as_uint(p.x) != __float_as_uint(p0) || __float_as_uint(p.y) != __float_as_uint(p1)) {
    bool nan_ok = (s.x != s.x) == (s0 != s0) && (s.y != s.y) == (s1 != s1) && (p.x != p.x) == (p0 != p0) && (p.y != p.y) == (p1 != p1);
    bool anynan = (s0 != s0) || (s1 != s1) || (p0 != p0) || (p1 != p1);
    if (!(anynan && nan_ok)) atomicAdd(bad, 1u);
  }
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2];
  for (int v = 0; v < 2; v++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (v == 0) k_scalar<<<blocks, threads>>>(out, 1.0f, iters); else k_packed<<<blocks, threads>>>(out, 1.0f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
    }
    double flops = (double)blocks * threads * iters * 2.0 * 2 * CH;
    printf("%s: %.3f ms, %.2f TFLOP/s fp32 (add+mul, no FMA)\n", v == 0 ? "scalar FADD/FMUL" : "packed FADD2/FMUL2", ms[v], flops / ms[v] * 1e-9);
  }
  printf("packed / scalar speed = %.2fx\n", ms[0] / ms[1]);
  const int n = 1 << 22;
  float *hx = (float*)malloc(n * 4), *hy = (float*)malloc(n * 4), *x, *y; unsigned* bad, hbad = 0;
  srand(7);
  for (int i = 0; i < n; i++) {
    unsigned a = ((unsigned)rand() << 16) ^ (unsigned)rand() ^ ((unsigned)rand() << 31), b = ((unsigned)rand() << 16) ^ (unsigned)rand() ^ ((unsigned)rand() << 31);
    if (i % 3 == 0) { float f = (float)(rand() % 2000 - 1000) * 1e-3f; memcpy(&a, &f, 4); }   // ordinary magnitudes too
    memcpy(hx + i, &a, 4); memcpy(hy + i, &b, 4);
  }
  cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  cudaMemcpy(x, hx, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(y, hy, n * 4, cudaMemcpyHostToDevice);
  k_check<<<n / 2 / 256, 256>>>(x, y, bad, n);
  cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
  printf("rounding check on %d random pairs (denormals, infinities, NaNs included): %u mismatches vs scalar add/mul\n", n / 2, hbad);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return hbad != 0;
}
