// microbenchmark of 16x16 base-case variants (one warp)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int tile_off(int r, int c) { return ((((c >> 5) << 3) + (r >> 3)) << 8) + (((c >> 2) & 7) << 5) + ((r & 7) << 2) + (c & 3); }

template <int VAR>
__global__ void __launch_bounds__(32) k_base(const double* A, double* out, long long* cyc, int reps) {
  __shared__ double sL[4096];
  __shared__ double sWd[1024];
  __shared__ double sBase[512];
  __shared__ double sCol[64];
  const int lane = threadIdx.x;
  for (int i = lane; i < 4096; i += 32) sL[i] = A[i];
  __syncwarp();
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    const int sb = rep & 3;
    const int base = 16 * sb, r = lane & 15;
    if (VAR == 0) {
      double row[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (lane < 16) row[c] = (c <= r) ? sL[tile_off(base + r, base + c)] : 0.0;
        else row[c] = (c == r) ? 1.0 : 0.0;
      }
      double d = __shfl_sync(0xffffffffu, row[0], 0);
      double rinv = rsqrt(d);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double lk = row[k] * rinv;
        row[k] = lk;
        double rinv_next = 0.0;
        if (k < 15) {
          const double dn = __shfl_sync(0xffffffffu, fma(-lk, lk, row[k + 1]), k + 1);
          rinv_next = rsqrt(dn);
        }
        double* col = sCol + ((k & 1) << 4);
        if (lane < 16) col[lane] = lk;
        __syncwarp();
#pragma unroll
        for (int c = k + 1; c < 16; ++c) row[c] = fma(-lk, col[c], row[c]);
        rinv = rinv_next;
      }
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c) sL[tile_off(base + r, base + c)] = (c <= r) ? row[c] : 0.0;
      } else {
#pragma unroll
        for (int cp = 0; cp < 16; ++cp)
          sWd[sb * 256 + ((((cp >> 3) << 2) + (r >> 2)) << 5) + ((cp & 7) << 2) + (r & 3)] = (cp >= r) ? row[cp] : 0.0;
      }
      __syncwarp();
    } else if (VAR == 1) {
      // smem-resident right-looking, rolled: S[c*32 + lane]
      for (int c = 0; c < 16; ++c) {
        double v;
        if (lane < 16) v = (c <= r) ? sL[tile_off(base + r, base + c)] : 0.0; else v = (c == r) ? 1.0 : 0.0;
        sBase[c * 32 + lane] = v;
      }
      __syncwarp();
#pragma unroll 1
      for (int k = 0; k < 16; ++k) {
        const double d = sBase[k * 32 + k];
        const double rinv = rsqrt(d);
        const double lk = sBase[k * 32 + lane] * rinv;
        __syncwarp();
        sBase[k * 32 + lane] = lk;
        __syncwarp();
#pragma unroll 5
        for (int c = k + 1; c < 16; ++c) sBase[c * 32 + lane] = fma(-lk, sBase[k * 32 + c], sBase[c * 32 + lane]);
        __syncwarp();
      }
      for (int c = 0; c < 16; ++c)
        if (lane < 16) sL[tile_off(base + r, base + c)] = (c <= r) ? sBase[c * 32 + lane] : 0.0;
      for (int q = 0; q < 8; ++q) {
        const int idx = lane + 32 * q;
        const int cp = (q >> 2) * 8 + (lane >> 2), kk = (q & 3) * 4 + (lane & 3);
        sWd[sb * 256 + idx] = (kk <= cp) ? sBase[cp * 32 + 16 + kk] : 0.0;
      }
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x] = (t1 - t0) / reps;
  out[blockIdx.x * 32 + lane] = sL[lane] + sWd[lane];
}
int main() {
  double* A; double* out; long long* cyc;
  cudaMallocManaged(&A, 4096 * 8); cudaMallocManaged(&out, 148 * 32 * 8); cudaMallocManaged(&cyc, 148 * 8);
  for (int i = 0; i < 4096; ++i) A[i] = 0.0;
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) { int o = ((((c >> 5) << 3) + (r >> 3)) << 8) + (((c >> 2) & 7) << 5) + ((r & 7) << 2) + (c & 3); A[o] = (r == c) ? 100.0 + r : 1.0 / (1 + r + c); }
  for (int pass = 0; pass < 2; ++pass) {
    k_base<0><<<148, 32>>>(A, out, cyc, 64); cudaDeviceSynchronize(); printf("VAR0 (registers, unrolled): %lld clk per base case\n", cyc[0]);
    k_base<1><<<148, 32>>>(A, out, cyc, 64); cudaDeviceSynchronize(); printf("VAR1 (smem rolled):          %lld clk per base case\n", cyc[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
