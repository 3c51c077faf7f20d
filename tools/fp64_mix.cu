// Are the FP64 tensor pipe (DMMA) and the FP64 vector pipe (DFMA/DADD) independent on B200?
// Even warps run a DMMA stream, odd warps a DFMA (or |a-b| DADD) stream; compare with each alone.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode bit0: even warps DMMA ; bit1: odd warps DFMA ; bit2: odd warps DADD-absdiff
__global__ void k_mix(double* out, int iters, int mode, int all_same) {
  const int warp = threadIdx.x >> 5;
  const bool tensor_role = all_same ? (mode & 1) : ((warp & 1) == 0);
  double s = 0;
  if (tensor_role) {
    if (!(mode & 1)) return;
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) mma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  } else if (mode & 2) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters * 8; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], 1.0000001, 1e-9);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
  } else if (mode & 4) {
    double acc[16], x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { acc[i] = 0; x[i] = threadIdx.x * 0.001 + i; }
    double y = 0.5;
    for (int it = 0; it < iters * 4; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] += fabs(x[i] - y);
      y += 1e-9;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
  } else return;
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount, iters = 8000, thr = 256, grid = nsm * 2;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int mode, int all_same) {
    k_mix<<<grid, thr>>>(out, iters, mode, all_same); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k_mix<<<grid, thr>>>(out, iters, mode, all_same); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
    return (double)best;
  };
  const double warps_half = (double)grid * thr / 32 / 2;
  const double mma_flops = 2.0 * 256 * 16 * iters * warps_half;         // even warps
  const double fma_flops = 2.0 * 16 * (iters * 8.0) * warps_half * 32;   // odd warps
  const double add_instr = 2.0 * 16 * (iters * 4.0) * warps_half * 32;
  double t;
  t = run(1, 0); printf("{\"dmma_only_half_warps_tflops\": %.2f,\n", mma_flops / t / 1e9);
  t = run(2, 0); printf(" \"dfma_only_half_warps_tflops\": %.2f,\n", fma_flops / t / 1e9);
  t = run(4, 0); printf(" \"dadd_only_half_warps_tinstr\": %.2f,\n", add_instr / t / 1e9);
  t = run(3, 0); printf(" \"mix_dmma_dfma_ms\": %.3f, \"mix_dmma_tflops\": %.2f, \"mix_dfma_tflops\": %.2f, \"mix_sum_tflops\": %.2f,\n", t, mma_flops / t / 1e9, fma_flops / t / 1e9, (mma_flops + fma_flops) / t / 1e9);
  t = run(5, 0); printf(" \"mix_dmma_dadd_ms\": %.3f, \"mix2_dmma_tflops\": %.2f, \"mix2_dadd_tinstr\": %.2f,\n", t, mma_flops / t / 1e9, add_instr / t / 1e9);
  double t1 = run(1, 0), t2 = run(2, 0), t3 = run(4, 0);
  printf(" \"alone_ms\": [%.3f, %.3f, %.3f], \"done\": 1}\n", t1, t2, t3);
  return 0;
}
