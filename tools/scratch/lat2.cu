// single-warp latency / issue-rate microbenchmarks (calibration for the band kernel's cost model)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rsq_seed(double d) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); return y; }
__global__ void k(long long* out, double x, int iters) {
  __shared__ double sh[1024];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = x + i;
  __syncthreads();
  long long t0, t1;
  double a = x + lane, b = 1.0 + 1e-9 * lane;
  {  // dependent DMMA chain
    double c0 = 0, c1 = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); }
    t1 = clock64(); if (threadIdx.x == 0) out[0] = t1 - t0; if (c0 + c1 == 1.2345) out[15] = 1;
  }
  {  // 2 independent chains
    double c[2][2] = {};
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { for (int u = 0; u < 2; ++u) for (int q = 0; q < 2; ++q) dmma(c[q][0], c[q][1], a, b); }
    t1 = clock64(); if (threadIdx.x == 0) out[1] = t1 - t0; if (c[0][0] + c[1][1] == 1.2345) out[15] = 1;
  }
  {  // 4 independent chains
    double c[4][2] = {};
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { for (int q = 0; q < 4; ++q) dmma(c[q][0], c[q][1], a, b); }
    t1 = clock64(); if (threadIdx.x == 0) out[2] = t1 - t0; if (c[0][0] + c[1][1] + c[2][0] + c[3][1] == 1.2345) out[15] = 1;
  }
  {  // 8 independent chains
    double c[8][2] = {};
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { for (int q = 0; q < 8; ++q) dmma(c[q][0], c[q][1], a, b); }
    t1 = clock64(); if (threadIdx.x == 0) out[3] = t1 - t0; double s = 0; for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1]; if (s == 1.2345) out[15] = 1;
  }
  {  // DFMA dependent
    double f = a;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { f = fma(f, 0.999999, 1e-9); f = fma(f, 0.999999, 1e-9); f = fma(f, 0.999999, 1e-9); f = fma(f, 0.999999, 1e-9); }
    t1 = clock64(); if (threadIdx.x == 0) out[4] = t1 - t0; if (f == 1.2345) out[15] = 1;
  }
  {  // DFMA 8 independent
    double f[8]; for (int q = 0; q < 8; ++q) f[q] = a + q;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { for (int q = 0; q < 8; ++q) f[q] = fma(f[q], 0.999999, 1e-9); }
    t1 = clock64(); if (threadIdx.x == 0) out[5] = t1 - t0; double s = 0; for (int q = 0; q < 8; ++q) s += f[q]; if (s == 1.2345) out[15] = 1;
  }
  {  // LDS.64 dependent chain (pointer chasing through values)
    int idx = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { double v = sh[idx]; idx = ((int)v + lane) & 1023; v = sh[idx]; idx = ((int)v + lane) & 1023; }
    t1 = clock64(); if (threadIdx.x == 0) out[6] = t1 - t0; if (idx == 12345) out[15] = 1;
  }
  {  // rsqrt seed + 2 newton chain
    double f = 2.0 + lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int u = 0; u < 2; ++u) { double y = rsq_seed(f); const double h = 0.5 * f; double e = fma(-h * y, y, 0.5); y = fma(y, e, y); e = fma(-h * y, y, 0.5); y = fma(y, e, y); f = y + 1.5; }
    }
    t1 = clock64(); if (threadIdx.x == 0) out[7] = t1 - t0; if (f == 1.2345) out[15] = 1;
  }
  {  // rsqrt seed only chain
    double f = 2.0 + lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { f = rsq_seed(f) + 1.5; f = rsq_seed(f) + 1.5; }
    t1 = clock64(); if (threadIdx.x == 0) out[8] = t1 - t0; if (f == 1.2345) out[15] = 1;
  }
  {  // shfl double chain
    double c = a;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { c = __shfl_sync(0xffffffffu, c, (lane + 1) & 31); c = __shfl_sync(0xffffffffu, c, (lane + 3) & 31); }
    t1 = clock64(); if (threadIdx.x == 0) out[9] = t1 - t0; if (c == 1.2345) out[15] = 1;
  }
  {  // STS -> syncwarp -> LDS round trip chain
    double c = a;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { sh[lane] = c; __syncwarp(); c = sh[(lane + 1) & 31] + 1.0; __syncwarp(); sh[lane] = c; __syncwarp(); c = sh[(lane + 5) & 31] + 1.0; __syncwarp(); }
    t1 = clock64(); if (threadIdx.x == 0) out[10] = t1 - t0; if (c == 1.2345) out[15] = 1;
  }
  {  // DMUL dependent
    double f = 1.0 + 1e-9 * lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { f = f * 1.0000001; f = f * 0.9999999; f = f * 1.0000001; f = f * 0.9999999; }
    t1 = clock64(); if (threadIdx.x == 0) out[11] = t1 - t0; if (f == 1.2345) out[15] = 1;
  }
}
int main() {
  long long* o; cudaMallocManaged(&o, 128); int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(o, 1.0, iters); cudaDeviceSynchronize(); }
  printf("single warp on an idle SM\n");
  printf("DMMA dependent chain:        %.1f clk per DMMA\n", o[0] / (4.0 * iters));
  printf("DMMA 2 independent chains:   %.1f clk per DMMA\n", o[1] / (4.0 * iters));
  printf("DMMA 4 independent chains:   %.1f clk per DMMA\n", o[2] / (4.0 * iters));
  printf("DMMA 8 independent chains:   %.1f clk per DMMA\n", o[3] / (8.0 * iters));
  printf("DFMA dependent:              %.1f clk\n", o[4] / (4.0 * iters));
  printf("DFMA 8 independent:          %.1f clk per DFMA\n", o[5] / (8.0 * iters));
  printf("LDS.64 dependent (+cvt,add): %.1f clk\n", o[6] / (2.0 * iters));
  printf("rsqrt seed + 2 newton + add: %.1f clk\n", o[7] / (2.0 * iters));
  printf("rsqrt seed + add:            %.1f clk\n", o[8] / (2.0 * iters));
  printf("shfl(double) chain:          %.1f clk\n", o[9] / (2.0 * iters));
  printf("STS->syncwarp->LDS->add:     %.1f clk\n", o[10] / (2.0 * iters));
  printf("DMUL dependent:              %.1f clk\n", o[11] / (4.0 * iters));
  // 4 warps per SM (one per scheduler) and 16 warps per SM: DMMA 8-chain rate
  for (int nt : {128, 512}) { k<<<1, nt>>>(o, 1.0, iters); cudaDeviceSynchronize();
    printf("%d threads in the CTA: DMMA dep %.1f, 8 chains %.1f clk per DMMA (warp 0); DFMA dep %.1f\n", nt, o[0] / (4.0 * iters), o[3] / (8.0 * iters), o[4] / (4.0 * iters)); }
  return 0;
}
