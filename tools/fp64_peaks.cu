// Micro-benchmarks: FP64 DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16) and DFMA issue
// rates on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_884(double* out, int iters) {
  double c[ILP][2];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_16816(double* out, int iters) {
  double c[ILP][4];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma(double* out, int iters) {
  double c[ILP];
  for (int i = 0; i < ILP; ++i) c[i] = i;
  double a = 1.0000001, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma(float* out, int iters) {
  float c[ILP];
  for (int i = 0; i < ILP; ++i) c[i] = i;
  float a = 1.0000001f, b = 1e-9f * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s SMs=%d\n", p.name, sms);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int blocks = sms * (warps * 32 / threads);
    float ms = time_ms([&] { dmma_884<8><<<blocks, threads>>>(out, iters); });
    double flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * (blocks * threads / 32);
    printf("DMMA m8n8k4   warps/SM=%2d  %8.2f TFLOP/s\n", warps, flops / ms / 1e9);
    ms = time_ms([&] { dmma_16816<4><<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 16 * 8 * 16 * 4 * (double)iters * (blocks * threads / 32);
    printf("DMMA m16n8k16 warps/SM=%2d  %8.2f TFLOP/s\n", warps, flops / ms / 1e9);
    ms = time_ms([&] { dfma<8><<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 8 * (double)iters * blocks * threads;
    printf("DFMA          warps/SM=%2d  %8.2f TFLOP/s\n", warps, flops / ms / 1e9);
    ms = time_ms([&] { ffma<8><<<blocks, threads>>>((float*)out, iters); });
    printf("FFMA          warps/SM=%2d  %8.2f TFLOP/s\n", warps, flops / ms / 1e9);
  }
  return 0;
}
