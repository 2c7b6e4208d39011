// Microbenchmark: unfused fp32 mul+add throughput, scalar (FMUL+FADD) vs packed (FMUL2+FADD2, sm_100 f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int ILP>
__global__ void k_scalar(float* out, float a, float b) {
  float x[2 * ILP];
  for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) x[i] = x[i] * a + b;  // FMUL + FADD (fmad=false)
  }
  float s = 0;
  for (int i = 0; i < 2 * ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_packed(float* out, float a, float b) {
  float2 x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = __fadd2_rn(__fmul2_rn(x[i], aa), bb);  // FMUL2 + FADD2
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int grid = 148 * 8, block = 256;
  const double ops = (double)grid * block * ITERS;  // per (mul+add) pair per value
  float t;
  t = timeit([&] { k_scalar<2><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("scalar ILP4 values/thread: %.3f ms  %.2f T(mul+add pairs)/s\n", t, ops * 4 / t / 1e9);
  t = timeit([&] { k_packed<2><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("packed ILP4 values/thread: %.3f ms  %.2f T(mul+add pairs)/s\n", t, ops * 4 / t / 1e9);
  t = timeit([&] { k_scalar<4><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("scalar ILP8 values/thread: %.3f ms  %.2f T(mul+add pairs)/s\n", t, ops * 8 / t / 1e9);
  t = timeit([&] { k_packed<4><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("packed ILP8 values/thread: %.3f ms  %.2f T(mul+add pairs)/s\n", t, ops * 8 / t / 1e9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
