// Instruction-rate probe for a sign-mask formulation of EMD: per element one FP32 subtract + one funnel shift
// (collecting the sign bit), versus the FP64 formulation (DADD + DADD|.|).  Which issues faster on B200?
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void k_fp64(double* out, int iters, double y0) {
  double acc[NACC], x[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i] = 0; x[i] = threadIdx.x * 0.001 + i; }
  double y = y0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] += fabs(x[i] - y);
    y += 1e-9;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_mask(double* out, int iters, float y0) {
  unsigned mask[NACC];
  float x[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { mask[i] = 0; x[i] = threadIdx.x * 0.001f + i * 0.37f; }
  float y = y0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      const float d = x[i] - y;
      mask[i] = __funnelshift_l(__float_as_uint(d), mask[i], 1);
    }
    y += 1e-3f;
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s ^= mask[i];
  if (s == 0x12345678u) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount, iters = 20000;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int thr : {128, 256, 512}) {
    const int grid = nsm * 2;
    auto time = [&](auto f) { f(); CK(cudaDeviceSynchronize()); float best = 1e30f; for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); } return (double)best; };
    const double elems = 64.0 * iters * grid * thr;
    double t64 = time([&] { k_fp64<64><<<grid, thr>>>(out, iters, 0.5); });
    double tm = time([&] { k_mask<64><<<grid, thr>>>(out, iters, 0.5f); });
    printf("threads %d: fp64 absdiff %.2f Telem/s, fp32 sub + funnel-shift %.2f Telem/s, ratio %.2f\n", thr, elems / t64 / 1e9, elems / tm / 1e9, t64 / tm);
  }
  return 0;
}
