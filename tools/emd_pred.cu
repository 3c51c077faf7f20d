// Does a predicated-off DADD occupy the FP64 pipe?  EMD could then be  2*sum max(d,0) - sum d  with
// d = a - b (DADD), sign test on the integer pipe, and a predicated accumulate.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>   // 0: acc += |x-y| ; 1: if (hi(d) >= 0) acc += d  (integer sign test)
__global__ void k(double* out, int iters, double y0, double step) {
  double acc[16], x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { acc[i] = 0; x[i] = (threadIdx.x & 1 ? 0.75 : 0.25) + i * 1e-3; }
  double y = y0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (MODE == 0) acc[i] += fabs(x[i] - y);
      else {
        const double d = x[i] - y;
        if (__double2hiint(d) >= 0) acc[i] += d;
      }
    }
    y += step;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount, iters = 40000, thr = 256, grid = nsm * 2;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto f) { f(); CK(cudaDeviceSynchronize()); float best = 1e30f; for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); } return (double)best; };
  const double elems = 16.0 * iters * grid * thr;
  double t0 = time([&] { k<0><<<grid, thr>>>(out, iters, 0.5, 1e-12); });
  // y = 0.5: odd threads (x=0.75) always positive, even threads (x=0.25) always negative -> per-lane 50 % taken, warp-divergent predicate
  double t1 = time([&] { k<1><<<grid, thr>>>(out, iters, 0.5, 1e-12); });
  double t2 = time([&] { k<1><<<grid, thr>>>(out, iters, 0.0, 1e-12); });   // always taken
  double t3 = time([&] { k<1><<<grid, thr>>>(out, iters, 1.0, 1e-12); });   // never taken
  printf("absdiff %.2f Telem/s | predicated: half lanes %.2f, always %.2f, never %.2f Telem/s\n", elems / t0 / 1e9, elems / t1 / 1e9, elems / t2 / 1e9, elems / t3 / 1e9);
  return 0;
}
