// Can the INT pipe take a share of the EMD inner loop while the FP64 pipe is saturated?  (lead 1)
// EMD = sum_k |U_k - V_k|: on the FP64 pipe two DADDs per element (the kernel runs at 92 % of that issue rate).  CDF
// values live in [0,1], so they can also be held as 64-bit fixed point (2^-48 resolution): |a-b| and the running sum
// are then exact integer operations on the INT pipe, which the EMD kernel leaves idle.  Each thread processes, per
// iteration, NF elements in FP64 and NI elements in int64, all register-resident (no memory traffic), with the EMD
// kernel's launch shape (2 CTAs of 128 threads per SM) and with 4x more warps.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NF, int NI, int IMODE>   // IMODE 0: acc += |x - y| ; 1: acc += min(x, y)  (sum|a-b| = sum a + sum b - 2 sum min)
__global__ void k(double* out, int iters, double y0, double step, long long yi0, long long stepi) {
  double accf[NF > 0 ? NF : 1], xf[NF > 0 ? NF : 1];
  long long acci[NI > 0 ? NI : 1], xi[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < NF; i++) { accf[i] = 0; xf[i] = (threadIdx.x & 1 ? 0.75 : 0.25) + i * 1e-3; }
#pragma unroll
  for (int i = 0; i < NI; i++) { acci[i] = 0; xi[i] = ((long long)((threadIdx.x & 1) ? 3 : 1) << 46) + i * 1000003ll + threadIdx.x; }
  double y = y0;
  long long yi = yi0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NF; i++) accf[i] += fabs(xf[i] - y);
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (IMODE == 0) {
        const long long d = xi[i] - yi;
        const long long s = d >> 63;
        acci[i] += (d ^ s) - s;
      } else {
        acci[i] += (long long)min((unsigned long long)xi[i], (unsigned long long)yi);
      }
    }
    y += step;
    yi += stepi;
  }
  double s = 0;
  long long si = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += accf[i];
#pragma unroll
  for (int i = 0; i < NI; i++) si += acci[i];
  if (s == 12345.678 || si == 0x123456789abcll) out[0] = s + (double)si;
}

cudaEvent_t e0, e1;
template <int NF, int NI, int IMODE>
void run(double* out, int grid, int thr, int iters, const char* label) {
  auto launch = [&] { k<NF, NI, IMODE><<<grid, thr>>>(out, iters, 0.5, 1e-12, 1ll << 47, 12345ll); };
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  const double elems = (double)(NF + NI) * iters * grid * thr;
  printf("{\"shape\": \"%s\", \"int_op\": \"%s\", \"fp64_elems\": %d, \"int_elems\": %d, \"ms\": %.3f, \"telem_per_s\": %.3f, \"fp64_telem_per_s\": %.3f}\n",
         label, IMODE ? "min" : "absdiff", NF, NI, best, elems / best / 1e9, (double)NF * iters * grid * thr / best / 1e9);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount, iters = 20000;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int shape = 0; shape < 2; ++shape) {
    const int thr = shape == 0 ? 128 : 512, grid = nsm * 2;
    const char* label = shape == 0 ? "2x128 threads per SM (EMD kernel)" : "2x512 threads per SM";
    run<16, 0, 0>(out, grid, thr, iters, label);
    run<16, 1, 0>(out, grid, thr, iters, label);
    run<16, 2, 0>(out, grid, thr, iters, label);
    run<16, 3, 0>(out, grid, thr, iters, label);
    run<16, 4, 0>(out, grid, thr, iters, label);
    run<16, 6, 0>(out, grid, thr, iters, label);
    run<13, 3, 0>(out, grid, thr, iters, label);
    run<0, 16, 0>(out, grid, thr, iters, label);
    run<16, 2, 1>(out, grid, thr, iters, label);
    run<16, 3, 1>(out, grid, thr, iters, label);
    run<16, 4, 1>(out, grid, thr, iters, label);
    run<16, 6, 1>(out, grid, thr, iters, label);
    run<16, 8, 1>(out, grid, thr, iters, label);
    run<12, 4, 1>(out, grid, thr, iters, label);
    run<0, 16, 1>(out, grid, thr, iters, label);
  }
  return 0;
}
