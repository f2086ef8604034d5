// Legacy warp-level mma.sync rates on sm_100a (tf32 m16n8k8, bf16 m16n8k16).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void tf32_1688(float* out, int iters) {
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x, 7};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void bf16_16816(float* out, int iters) {
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x, 7};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 2);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32, blocks = sms;
    float ms = time_ms([&] { tf32_1688<8><<<blocks, threads>>>(out, iters); });
    double flops = 2.0 * 16 * 8 * 8 * 8 * (double)iters * (blocks * warps);
    printf("mma.sync tf32 m16n8k8   warps/SM=%2d %9.1f TFLOP/s\n", warps, flops / ms / 1e9);
    ms = time_ms([&] { bf16_16816<8><<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 16 * 8 * 16 * 8 * (double)iters * (blocks * warps);
    printf("mma.sync bf16 m16n8k16  warps/SM=%2d %9.1f TFLOP/s\n", warps, flops / ms / 1e9);
  }
  return 0;
}
