// FP64 pipe microbenchmarks.  MEASURED_PEAKS.json carries HBM and bf16 numbers only, so the
// FP64 roofline denominators (vector DFMA, tensor DMMA) are measured on the box by these.
#include <stdio.h>
#include <stdlib.h>

#include "tb_common.cuh"
#include "tb_blocks.cuh"

namespace {

__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters, double x) {
  double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
  const double m = x, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dmma884(double* out, int iters, double x) {
  double c[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) c[q][0] = c[q][1] = 0.0;
  const double a = x, b = 1.0 + 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dmma1688(double* out, int iters, double x) {
  double c[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.0;
  const double a = x, b = 1.0 + 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3])
                     : "d"(a), "d"(a), "d"(a), "d"(a), "d"(b), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += c[q][0] + c[q][1] + c[q][2] + c[q][3];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// accuracy probe of the pivot reciprocal square root (tb_blocks.cuh: rsqrt_pos): values log-uniform over
// [1e-290, 1e300] (the range the factorisation accepts), error against 1 / sqrt(d) (correctly rounded steps)
__global__ void k_rsqrt_probe(int n, double* worst) {
  double w = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double t = (i + 0.5) / n;
    const double d = exp10(-290.0 + 590.0 * t) * (1.0 + 0.37 * t);
    const double got = tbblk::rsqrt_pos(d), want = 1.0 / sqrt(d);
    w = fmax(w, fabs(got - want) / want);
  }
  for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(worst), (unsigned long long)__double_as_longlong(w));
}

}  // namespace

namespace {
// TbDivisor against __ddiv_rn: numerators and divisors spread over magnitudes and significand patterns (random bits,
// all-ones / near-all-ones / power-of-two significands, tiny and huge exponents, zero numerators)
__global__ void k_div_probe(long long n, unsigned long long seed, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
    auto next = [&]() { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; return x * 0x2545F4914F6CDD1Dull; };
    const unsigned long long r0 = next(), r1 = next(), r2 = next();
    unsigned long long mb = r0 & 0xfffffffffffffull, ma = r1 & 0xfffffffffffffull;
    const int pat = (int)(r2 & 15);
    if (pat == 0) mb = 0xfffffffffffffull;                       // significand all ones
    if (pat == 1) mb = 0xffffffffffffeull;
    if (pat == 2) mb = 0;                                        // power of two
    if (pat == 3) mb = 1;
    if (pat == 4) ma = 0xfffffffffffffull;
    if (pat == 5) ma = mb;                                       // quotient a power of two
    int eb = 1023 + (int)((r2 >> 8) % 41) - 20, ea = 1023 + (int)((r2 >> 16) % 61) - 30;
    if (pat == 6) eb = 1 + (int)((r2 >> 24) % 40);               // tiny divisor
    if (pat == 7) eb = 2046 - (int)((r2 >> 24) % 40);            // huge divisor
    if (pat == 8) ea = 1 + (int)((r2 >> 24) % 40);
    if (pat == 9) ea = 2046 - (int)((r2 >> 24) % 40);
    double b = __longlong_as_double((long long)(((unsigned long long)eb << 52) | mb));
    double a = __longlong_as_double((long long)(((unsigned long long)ea << 52) | ma));
    if (pat == 10) a = 0.0;
    if (r2 & (1ull << 40)) a = -a;
    const TbDivisor d(b);
    const double got = d.div(a), want = __ddiv_rn(a, b);
    if (__double_as_longlong(got) != __double_as_longlong(want)) {
      ++bad;
      atomicAdd(mismatches + 1 + pat, 1ull);                     // (per operand pattern, for diagnosis)
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int tb_div_probe(int64_t n, uint64_t seed, uint64_t* mismatches) {
  if (!mismatches) return TB_ERR_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  unsigned long long* d = nullptr;
  TB_CUDA(cudaMalloc(&d, 17 * sizeof(unsigned long long)));
  TB_CUDA(cudaMemset(d, 0, 17 * sizeof(unsigned long long)));
  k_div_probe<<<148 * 8, 256>>>((long long)n, (unsigned long long)seed, d);
  tb_count_launch(1);
  unsigned long long h[17] = {};
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  *mismatches = (uint64_t)h[0];
  if (getenv("TB_DIV_PROBE_DEBUG"))
    for (int i = 0; i < 16; ++i) fprintf(stderr, "[tb_div_probe] pattern %2d: %llu mismatches\n", i, h[1 + i]);
  return (int)e;
}

extern "C" int tb_rsqrt_probe(int32_t n, double* max_rel_err) {
  if (!max_rel_err) return TB_ERR_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  double* d = nullptr;
  TB_CUDA(cudaMalloc(&d, sizeof(double)));
  TB_CUDA(cudaMemset(d, 0, sizeof(double)));
  k_rsqrt_probe<<<256, 256>>>(n, d);
  tb_count_launch(1);
  cudaError_t e = cudaMemcpy(max_rel_err, d, sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (int)e;
}

extern "C" int tb_fp64_peak(int32_t which, int32_t iters, double* tflops, float* ms_out) {
  if (!tflops) return TB_ERR_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (iters <= 0) iters = 4096;
  const int blocks = sms * 8, threads = 256;
  double* out = nullptr;
  TB_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double flops = 0.0;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (which == 0) {
      k_peak_dfma<<<blocks, threads>>>(out, iters, 0.999999);
      flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    } else if (which == 1) {
      k_peak_dmma884<<<blocks, threads>>>(out, iters, 0.5);
      flops = 512.0 * 32.0 * iters * (double)blocks * (threads / 32);
    } else {
      k_peak_dmma1688<<<blocks, threads>>>(out, iters, 0.5);
      flops = 2048.0 * 16.0 * iters * (double)blocks * (threads / 32);
    }
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(out); return (int)e; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  tb_count_launch(4);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return (int)cudaGetLastError();
}
