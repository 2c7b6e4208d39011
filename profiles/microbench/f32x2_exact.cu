// Are the packed f32x2 intrinsics exact and un-contracted under -fmad=false?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_exact f32x2_exact.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ float rf(uint32_t& s) {  // random float with moderate exponent
  uint32_t m = rnd(s) & 0x007fffffu, e = 120 + (rnd(s) >> 28), sg = rnd(s) & 0x80000000u;
  return __uint_as_float(sg | (e << 23) | m);
}
__global__ void k(unsigned long long* bad) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 1;
  unsigned long long b[6] = {0, 0, 0, 0, 0, 0};
  for (int it = 0; it < 20000; it++) {
    float a0 = rf(s), a1 = rf(s), b0 = rf(s), b1 = rf(s), c0 = rf(s), c1 = rf(s);
    float2 A = make_float2(a0, a1), B = make_float2(b0, b1), C = make_float2(c0, c1);
    float2 m = mul2(A, B);
    b[0] += (__float_as_uint(m.x) != __float_as_uint(__fmul_rn(a0, b0))) + (__float_as_uint(m.y) != __float_as_uint(__fmul_rn(a1, b1)));
    float2 ad = add2(A, B);
    b[1] += (__float_as_uint(ad.x) != __float_as_uint(__fadd_rn(a0, b0))) + (__float_as_uint(ad.y) != __float_as_uint(__fadd_rn(a1, b1)));
    float2 ma = add2(C, mul2(A, B));  // must be fl(c + fl(a*b)), not an fma
    float r0 = __fadd_rn(c0, __fmul_rn(a0, b0)), r1 = __fadd_rn(c1, __fmul_rn(a1, b1));
    b[2] += (__float_as_uint(ma.x) != __float_as_uint(r0)) + (__float_as_uint(ma.y) != __float_as_uint(r1));
    float2 sb = sub2(A, B);
    b[3] += (__float_as_uint(sb.x) != __float_as_uint(__fadd_rn(a0, -b0))) + (__float_as_uint(sb.y) != __float_as_uint(__fadd_rn(a1, -b1)));
    float2 ms = mul2(sub2(C, A), B);
    b[4] += (__float_as_uint(ms.x) != __float_as_uint(__fmul_rn(__fadd_rn(c0, -a0), b0))) + (__float_as_uint(ms.y) != __float_as_uint(__fmul_rn(__fadd_rn(c1, -a1), b1)));
    float2 d3 = add2(add2(mul2(A, A), mul2(B, B)), mul2(C, C));
    float e0 = __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(b0, b0)), __fmul_rn(c0, c0));
    float e1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)), __fmul_rn(c1, c1));
    b[5] += (__float_as_uint(d3.x) != __float_as_uint(e0)) + (__float_as_uint(d3.y) != __float_as_uint(e1));
  }
  for (int i = 0; i < 6; i++) if (b[i]) atomicAdd(bad + i, b[i]);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
  k<<<148, 256>>>(d);
  unsigned long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("mismatches: mul2 %llu add2 %llu add2(c,mul2) %llu sub2 %llu mul2(sub2) %llu dot %llu  (%s)\n", h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
