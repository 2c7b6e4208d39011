// Microbenchmark: issue rate of FFMA vs FFMA2 (sm_100 packed f32x2), alone and interleaved with integer ALU work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
template <int N>
__global__ void k_ffma(float* out, float a, float b) {
  float x[N];
  for (int i = 0; i < N; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < N; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int N>
__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[N];
  for (int i = 0; i < N; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
  for (int i = 0; i < N; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// N FFMA2 + N integer ops per iteration
template <int N>
__global__ void k_ffma2_alu(float* out, float a, float b, unsigned m) {
  float2 x[N];
  unsigned y[N];
  for (int i = 0; i < N; i++) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); y[i] = threadIdx.x + i; }
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = __ffma2_rn(x[i], aa, bb); y[i] = (y[i] ^ m) + (y[i] >> 3); }
  }
  float s = 0;
  for (int i = 0; i < N; i++) s += x[i].x + x[i].y + (float)y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int N>
__global__ void k_ffma_alu(float* out, float a, float b, unsigned m) {
  float x[2 * N];
  unsigned y[N];
  for (int i = 0; i < 2 * N; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < N; i++) y[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) { x[2 * i] = __fmaf_rn(x[2 * i], a, b); x[2 * i + 1] = __fmaf_rn(x[2 * i + 1], a, b); y[i] = (y[i] ^ m) + (y[i] >> 3); }
  }
  float s = 0;
  for (int i = 0; i < 2 * N; i++) s += x[i];
  for (int i = 0; i < N; i++) s += (float)y[i];
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
  const double thr = (double)grid * block * ITERS;
  float t;
  t = timeit([&] { k_ffma<8><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("FFMA  x8 : %.3f ms  %.2f Tfma/s  (%.2f warp-instr/clk/SMSP @1.965GHz)\n", t, thr * 8 / t / 1e9, thr * 8 / 32 / (t * 1e-3) / (148 * 4 * 1.965e9));
  t = timeit([&] { k_ffma2<4><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2 x4 : %.3f ms  %.2f Tfma/s  (%.2f warp-instr/clk/SMSP)\n", t, thr * 8 / t / 1e9, thr * 4 / 32 / (t * 1e-3) / (148 * 4 * 1.965e9));
  t = timeit([&] { k_ffma2<8><<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2 x8 : %.3f ms  %.2f Tfma/s  (%.2f warp-instr/clk/SMSP)\n", t, thr * 16 / t / 1e9, thr * 8 / 32 / (t * 1e-3) / (148 * 4 * 1.965e9));
  t = timeit([&] { k_ffma_alu<4><<<grid, block>>>(out, 1.0001f, 0.5f, 0x5bd1e995u); });
  printf("8 FFMA + 4x(3 ALU)  : %.3f ms\n", t);
  t = timeit([&] { k_ffma2_alu<4><<<grid, block>>>(out, 1.0001f, 0.5f, 0x5bd1e995u); });
  printf("4 FFMA2 + 4x(3 ALU) : %.3f ms\n", t);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
