// 16x16-block building blocks shared by the band kernels (tb_band.cu) and the fused small-system kernel
// (tb_dense16.cu): FP64 tensor-core MMA, the fragment-major block layout, the register-resident 16x16 base case.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tbblk {

constexpr int BT = 16;    // block order
constexpr int BE = 256;   // doubles per block, stored as [8-row block 2][k-slab 4][lane 32] (DMMA operand order)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// Prefetch loads as volatile asm: the compiler keeps them in program order relative to the (volatile) DMMAs instead
// of sinking them next to their first use, so the tensor work of a block column really covers their latency.
__device__ __forceinline__ int ldg_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// Block layout: 8 slabs (slab s = 4 * (row / 8) + k / 4) of 32 doubles; inside a slab element p = 4 * (row % 8) + k % 4
// sits at p, so an operand read (lane = p, all 32 elements of one slab) is conflict free.  An XOR swizzle of p
// (swz(s) = ((s & 1) << 3) | ((s & 2) << 1), plus bit 4 of p folded onto bit 1) removes the 2-way / 4-way conflicts of
// the accumulator-pair, column and W-row accesses (12.1 M -> 3.2 M conflict wavefronts per 1024 bar-942 systems,
// LSU pipe 69 % -> 55 %) but costs 7 % more instructions on the dependent chains and the kernel got 3 % slower
// (profiles/r01g notes): the band kernels are bound by chain latency, not by the LSU pipe, so the plain layout stays.
__host__ __device__ __forceinline__ constexpr int swz(int) { return 0; }
__host__ __device__ __forceinline__ constexpr int slab_off(int s, int p) { return (s << 5) + (p ^ swz(s)); }
__device__ __forceinline__ int lane_swz(int lane) { return lane; }
// operand fragment element of this lane in slab s; lsw = lane_swz(lane)
__device__ __forceinline__ int fo(int s, int lsw) { return (s << 5) + (lsw ^ swz(s)); }
__host__ __device__ __forceinline__ constexpr int b16_off(int r, int c) { return slab_off(((r >> 3) << 2) + (c >> 2), ((r & 7) << 2) + (c & 3)); }
// this lane's accumulator pair of 8x8 block (mb, nbp) inside a 16x16 block
__device__ __forceinline__ int cpair_off(int mb, int nbp, int lane) {
  return slab_off(mb * 4 + nbp * 2 + ((lane & 3) >> 1), ((lane >> 2) << 2) + ((lane & 1) << 1));
}


// 1/sqrt(d) for a normal positive d: hardware seed (MUFU.RSQ64H, relative error e0 ~ 2^-22) + one third-order step
// y (1 + e/2 + 3 e^2/8), e = 1 - d y^2: error O(e0^3), a dependent chain of four FP64 operations (two Newton steps
// need six).  The library rsqrt() costs 18 instructions with its special-case handling; the pivots here are checked
// positive beside the chain (a non-positive pivot yields NaNs that are discarded with the failure flag).
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  const double t = fma(0.375, e, 0.5);
  return fma(y * e, t, y);
}

// 16x16 diagonal block in shared memory (fragment layout): L_D L_D^T = P in registers, W = L_D^{-1} written back
// over P as a DMMA B operand.  One warp: lanes 0-15 hold the rows of the block, lanes 16-31 the rows of
// Z = L^{-T} (identity to start with).  The lane owning row k+1 forms the next pivot from its own registers,
// one shuffle broadcasts it and the reciprocal square root of column k+1 is issued before the trailing update
// of column k.  Returns 0 or the 1-based index (row0 + k + 1) of the first non-positive pivot.
__device__ __forceinline__ int factor_diag16(double* sBlk, double* sCol, int lane, int row0) {
  const int r = lane & 15;
  double row[16];
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) {
    const double v = sBlk[b16_off(r, cc)];
    row[cc] = lane < 16 ? (cc <= r ? v : 0.0) : (cc == r ? 1.0 : 0.0);
  }
  __syncwarp();
  int bad = 0;
  double d = __shfl_sync(0xffffffffu, row[0], 0);
  if (!(d > 1e-290)) bad = row0 + 1;
  double rinv = rsqrt_pos(d);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const double lk = row[k] * rinv;                        // lane k: d * rsqrt(d) = sqrt(d)
    row[k] = lk;
    double rinv_next = 0.0;
    if (k < 15) {
      const double dn = __shfl_sync(0xffffffffu, fma(-lk, lk, row[k + 1]), k + 1);
      rinv_next = rsqrt_pos(dn);
      if (!(dn > 1e-290) && !bad) bad = row0 + k + 2;       // same value in every lane; off the pivot chain
    }
    double* col = sCol + ((k & 1) << 4);
    if (lane < 16) col[lane] = lk;
    __syncwarp();
    const double2* col2 = reinterpret_cast<const double2*>(col);
#pragma unroll
    for (int p = (k + 1) >> 1; p < 8; ++p) {
      const double2 cv = col2[p];
      if (2 * p > k) row[2 * p] = fma(-lk, cv.x, row[2 * p]);
      row[2 * p + 1] = fma(-lk, cv.y, row[2 * p + 1]);
    }
    rinv = rinv_next;
  }
  if (!bad && lane >= 16) {
    // W[c'][kk] = Z[kk][c'] for kk <= c' (this lane: kk = r), as a DMMA B operand, over P
#pragma unroll
    for (int cp = 0; cp < 16; ++cp) sBlk[b16_off(cp, r)] = (cp >= r) ? row[cp] : 0.0;
  }
  return bad;
}


}  // namespace tbblk
