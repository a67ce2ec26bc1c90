// FP64 pipe microbenchmarks.  MEASURED_PEAKS.json carries HBM and bf16 numbers only, so the
// FP64 roofline denominators (vector DFMA, tensor DMMA) are measured on the box by these.
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
