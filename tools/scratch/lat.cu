#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(long long* out, double x, int iters) {
  double a = x + threadIdx.x; long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = fma(a, 0.999999, 1e-9); a = fma(a, 0.999999, 1e-9); a = fma(a, 0.999999, 1e-9); a = fma(a, 0.999999, 1e-9); }
  t1 = clock64(); if (threadIdx.x == 0) out[0] = (t1 - t0);
  // rsqrt chain
  double b = 2.0 + a * 1e-30;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { b = rsqrt(b) + 1.5; b = rsqrt(b) + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) out[1] = (t1 - t0);
  // shfl chain (double)
  double c = b;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { c = __shfl_sync(0xffffffffu, c, (threadIdx.x + 1) & 31); c = __shfl_sync(0xffffffffu, c, (threadIdx.x + 3) & 31); }
  t1 = clock64(); if (threadIdx.x == 0) out[2] = (t1 - t0);
  // 1/x chain
  double d = 1.7 + c * 1e-30;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { d = 1.0 / d + 0.5; d = 1.0 / d + 0.5; }
  t1 = clock64(); if (threadIdx.x == 0) out[3] = (t1 - t0);
  // sqrt chain
  double e = 1.7 + d * 1e-30;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { e = sqrt(e) + 0.5; e = sqrt(e) + 0.5; }
  t1 = clock64(); if (threadIdx.x == 0) out[4] = (t1 - t0);
  // float rsqrt + 2 newton in double
  double f = 1.7 + e * 1e-30;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int u = 0; u < 2; ++u) {
      double y = (double)rsqrtf((float)f);
      double h = 0.5 * f;
      y = y * fma(-h * y, y, 1.5);
      y = y * fma(-h * y, y, 1.5);
      f = y + 1.5;
    }
  }
  t1 = clock64(); if (threadIdx.x == 0) out[5] = (t1 - t0);
  if (a + b + c + d + e + f == 1234.5) out[7] = 1;
}
int main() {
  long long* o; cudaMallocManaged(&o, 64); int iters = 1000;
  k_lat<<<1, 32>>>(o, 1.0, iters); cudaDeviceSynchronize();
  k_lat<<<1, 32>>>(o, 1.0, iters); cudaDeviceSynchronize();
  printf("DFMA dependent latency: %.1f clk\n", o[0] / (4.0 * iters));
  printf("rsqrt(double)+add chain: %.1f clk\n", o[1] / (2.0 * iters));
  printf("shfl(double) chain: %.1f clk\n", o[2] / (2.0 * iters));
  printf("1/x + add chain: %.1f clk\n", o[3] / (2.0 * iters));
  printf("sqrt + add chain: %.1f clk\n", o[4] / (2.0 * iters));
  printf("rsqrtf seed + 2 newton + add: %.1f clk\n", o[5] / (2.0 * iters));
  return 0;
}
