// FP64 roofline denominators for B200 (sm_100a): DFMA / DADD issue rate, DMMA (mma.sync f64)
// shapes, and cuBLAS DGEMM 8192^3 burst + sustained. MEASURED_PEAKS.json has no FP64 figure
// (SURVEY.md section 7 hard part 2), so this tool writes the denominators the bench uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// EMD inner op: acc += |x - y|  -> 2 FP64 instructions (DADD, DADD with |.| modifier)
template <int NACC>
__global__ void k_dabsdiff(double* out, int iters, double y0) {
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

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NT>
__global__ void k_mma884(double* out, int iters) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) mma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
template <int NT, int KV>  // KV = 4, 8, 16
__global__ void k_mma16(double* out, int iters) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) {
      if (KV == 4) mma1684(c[i], a, b[0]);
      else if (KV == 8) mma1688(c[i], a, b);
      else mma16816(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
double time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, (double)ms);
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, nsm, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  const int iters = 20000;
  int cfg[][2] = {{1, 128}, {1, 256}, {1, 512}, {2, 256}, {1, 1024}, {2, 512}};  // blocks/SM, threads
  for (auto& c : cfg) {
    int grid = nsm * c[0], thr = c[1];
    double warps = (double)grid * thr / 32;
    double ms;
    ms = time_ms([&] { k_dfma<16><<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf(" \"dfma_b%d_t%d_tflops\": %.2f,\n", c[0], thr, 2.0 * 16 * iters * grid * thr / ms / 1e9);
    ms = time_ms([&] { k_dabsdiff<16><<<grid, thr>>>(out, iters, 0.5); });
    printf(" \"dabsdiff_b%d_t%d_tinstr\": %.2f,\n", c[0], thr, 2.0 * 16 * iters * grid * thr / ms / 1e9);
    ms = time_ms([&] { k_mma884<16><<<grid, thr>>>(out, iters / 4); });
    printf(" \"dmma884_b%d_t%d_tflops\": %.2f,\n", c[0], thr, 2.0 * 256 * 16 * (iters / 4) * warps / ms / 1e9);
    ms = time_ms([&] { k_mma16<8, 4><<<grid, thr>>>(out, iters / 4); });
    printf(" \"dmma1684_b%d_t%d_tflops\": %.2f,\n", c[0], thr, 2.0 * 512 * 8 * (iters / 4) * warps / ms / 1e9);
    ms = time_ms([&] { k_mma16<8, 8><<<grid, thr>>>(out, iters / 8); });
    printf(" \"dmma1688_b%d_t%d_tflops\": %.2f,\n", c[0], thr, 2.0 * 1024 * 8 * (iters / 8) * warps / ms / 1e9);
    ms = time_ms([&] { k_mma16<8, 16><<<grid, thr>>>(out, iters / 16); });
    printf(" \"dmma16816_b%d_t%d_tflops\": %.2f,\n", c[0], thr, 2.0 * 2048 * 8 * (iters / 16) * warps / ms / 1e9);
  }
  // cuBLAS DGEMM 8192^3
  {
    const int n = 8192;
    double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> h((size_t)n * n);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)rand() / RAND_MAX;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    cublasHandle_t hd; cublasCreate(&hd);
    double al = 1.0, be = 0.0;
    auto gemm = [&] { cublasDgemm(hd, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); };
    double ms = time_ms(gemm, 5);
    printf(" \"cublas_dgemm_8192_burst_tflops\": %.2f,\n", 2.0 * n * n * (double)n / ms / 1e9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int reps = std::max(3, (int)(4000.0 / ms));
    cudaEventRecord(e0); for (int i = 0; i < reps; i++) gemm(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float tot; cudaEventElapsedTime(&tot, e0, e1);
    printf(" \"cublas_dgemm_8192_sustained_tflops\": %.2f, \"sustained_reps\": %d,\n", 2.0 * n * n * (double)n * reps / tot / 1e9, reps);
  }
  // sustained DFMA / DMMA (4 s each) to see the power-capped rate
  {
    int grid = nsm * 2, thr = 256; double warps = (double)grid * thr / 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double ms1 = time_ms([&] { k_dfma<16><<<grid, thr>>>(out, iters, 1.0000001, 1e-9); }, 1);
    int reps = std::max(3, (int)(3000.0 / ms1));
    cudaEventRecord(e0); for (int i = 0; i < reps; i++) k_dfma<16><<<grid, thr>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float tot; cudaEventElapsedTime(&tot, e0, e1);
    printf(" \"dfma_sustained_tflops\": %.2f,\n", 2.0 * 16 * iters * grid * thr * (double)reps / tot / 1e9);
    ms1 = time_ms([&] { k_mma884<16><<<grid, thr>>>(out, iters / 4); }, 1);
    reps = std::max(3, (int)(3000.0 / ms1));
    cudaEventRecord(e0); for (int i = 0; i < reps; i++) k_mma884<16><<<grid, thr>>>(out, iters / 4);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tot, e0, e1);
    printf(" \"dmma884_sustained_tflops\": %.2f,\n", 2.0 * 256 * 16 * (iters / 4) * warps * (double)reps / tot / 1e9);
  }
  printf(" \"done\": 1}\n");
  return 0;
}
