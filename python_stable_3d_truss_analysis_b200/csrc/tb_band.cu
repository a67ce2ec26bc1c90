// Band path: block-band Cholesky for systems whose reduced stiffness matrix has a narrow
// envelope (bar-942: n = 696, half-bandwidth 56).  The dense/tiled pipeline of tb_large.cu spends
// its time on structurally zero 64x64 tiles and on latency between them; here the matrix is viewed
// as a band of 16x16 blocks (NB sub-diagonal blocks per block column) and ONE small CTA (4 warps)
// factorises one system with the active window of the band resident in shared memory:
//
//   per block column c:   assemble K blocks (c..c+NB, c) from the scatter map's values
//                         update     P(R,c) -= sum_{d=1..NB} L(R,c-d) L(c,c-d)^T           (DMMA m8n8k4)
//                         factor     16x16 diagonal block + its inverse W (one warp, in registers)
//                         solve      L(R,c) = P(R,c) W^T for the NB blocks below            (DMMA)
//                         forward    y_c = W (f_c - sum_d L(c,c-d) y_{c-d})
//   then block back-substitution  u_c = W_c^T (y_c - sum_{rb} L(c+rb,c)^T u_{c+rb}).
//
// Only (NB+1)(NB+2)/2 blocks are alive at any time (block (p+e, p) lives on "diagonal e", which needs
// a ring of e+1 slots), 30 KB for NB = 4, so seven CTAs share an SM and hide each other's pivot
// latency.  HBM sees the K values once, the off-diagonal L blocks and the 16x16 inverses once out and
// once back in (back-substitution), and u.  Replaces np.linalg.solve (slientruss3d/truss.py:343) for
// this class of systems; results equal the dense factorisation's (zeros are skipped, nothing else).
#include <math.h>

#include "tb_common.cuh"

namespace {

constexpr int BT = 16;    // block order
constexpr int BE = 256;   // doubles per block, stored as [8-row block 2][k-slab 4][lane 32] (DMMA operand order)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int b16_off(int r, int c) { return ((((r >> 3) << 2) + (c >> 2)) << 5) + ((r & 7) << 2) + (c & 3); }
// ring slot (in doubles) of block (row block p+e, column block p): diagonal e keeps e+1 blocks alive
__device__ __forceinline__ int slot(int e, int p) { return ((e * (e + 1)) / 2 + p % (e + 1)) * BE; }
// this lane's accumulator pair of 8x8 block (mb, nbp) inside a 16x16 block
__device__ __forceinline__ int cpair_off(int mb, int nbp, int lane) {
  return ((mb * 4 + nbp * 2 + ((lane & 3) >> 1)) << 5) + ((lane >> 2) << 2) + ((lane & 1) << 1);
}

__global__ void __launch_bounds__(TB_BAND_THREADS, 7) k_band(const LargeArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int NB = a.NB, ncol = a.nb16;
  const int nring = (NB + 1) * (NB + 2) / 2;
  double* sRing = sm;                          // alive blocks of the band (factorisation) | staging (back substitution)
  double* sY = sRing + nring * BE;             // ring of the last NB+1 blocks of y (then u), 16 each
  double* sT = sY + (NB + 1) * BT;             // [16] right-hand side of the current block
  double* sCol = sT + BT;                      // [32] base case: eliminated column, double buffered
  int* sFlag = (int*)(sCol + 32);
  double* sRed = sRing + (NB + 1) * BE;        // alias, back substitution only: [8][16] partial sums

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (a.status[b] != 0) continue;            // input problem flagged by k_geom (uniform)
    const double* kvs = a.kv + (int64_t)b * a.nnz;
    const double* fsys = a.force + b * a.force_stride;
    double* ysys = a.y + (int64_t)b * a.n_pad;
    double* Lb = a.L + (int64_t)b * ncol * (NB + 1) * BE;   // block (c+rb, c) at (c*(NB+1)+rb)*256, rb >= 1
    double* Wb = a.wd + (int64_t)b * ncol * BE;
    if (tid == 0) *sFlag = 0;
    int fail = 0;

    for (int c = 0; c < ncol; ++c) {
      // ---------------- A: assemble the K blocks of block column c into their ring slots
      for (int e = 0; e <= NB; ++e) {
        double* blk = sRing + slot(e, c);
        for (int i = tid; i < BE; i += TB_BAND_THREADS) blk[i] = 0.0;
      }
      __syncthreads();
      {
        const int q0 = a.b16_ptr[c], q1 = a.b16_ptr[c + 1];
        for (int q = q0 + tid; q < q1; q += TB_BAND_THREADS) {
          const int pos = a.b16_pos[q];        // e << 8 | offset inside the block
          sRing[slot(pos >> 8, c) + (pos & 255)] += kvs[q];
        }
        if (tid < BT) {
          const int row = c * BT + tid;
          if (row >= a.n) sRing[slot(0, c) + b16_off(tid, tid)] += 1.0;   // identity on the padded diagonal
        }
      }
      __syncthreads();

      // ---------------- B: left-looking update with the previous NB block columns (DMMA)
      for (int t = warp; t < 2 * (NB + 1); t += TB_BAND_THREADS / 32) {
        const int rb = t >> 1, mb = t & 1;
        double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, e0[2] = {0.0, 0.0}, e1[2] = {0.0, 0.0};
        for (int d = 1; d <= NB - rb && d <= c; ++d) {
          const double* A = sRing + slot(rb + d, c - d);   // L(c+rb, c-d)
          const double* Bm = sRing + slot(d, c - d);       // L(c,    c-d)
#pragma unroll
          for (int ks = 0; ks < 4; ks += 2) {
            const double av0 = A[((mb * 4 + ks) << 5) + lane], av1 = A[((mb * 4 + ks + 1) << 5) + lane];
            dmma(c0[0], c0[1], av0, Bm[(ks << 5) + lane]);
            dmma(c1[0], c1[1], av0, Bm[((4 + ks) << 5) + lane]);
            dmma(e0[0], e0[1], av1, Bm[((ks + 1) << 5) + lane]);
            dmma(e1[0], e1[1], av1, Bm[((4 + ks + 1) << 5) + lane]);
          }
        }
        double* tgt = sRing + slot(rb, c);
        double2* p0 = reinterpret_cast<double2*>(tgt + cpair_off(mb, 0, lane));
        double2* p1 = reinterpret_cast<double2*>(tgt + cpair_off(mb, 1, lane));
        double2 v0 = *p0, v1 = *p1;
        v0.x -= c0[0] + e0[0]; v0.y -= c0[1] + e0[1];
        v1.x -= c1[0] + e1[0]; v1.y -= c1[1] + e1[1];
        *p0 = v0; *p1 = v1;
      }
      // right-hand side of this block: f_c - sum_d L(c,c-d) y_{c-d}   (last warp: it has the fewest update tasks)
      if (warp == TB_BAND_THREADS / 32 - 1) {
        const int row = lane & 15, hh = lane >> 4;
        double tacc = 0.0;
        for (int d = 1; d <= NB && d <= c; ++d) {
          const double* Bm = sRing + slot(d, c - d);
          const double* yv = sY + ((c - d) % (NB + 1)) * BT;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) tacc = fma(Bm[b16_off(row, 2 * kk + hh)], yv[2 * kk + hh], tacc);
        }
        tacc += __shfl_xor_sync(0xffffffffu, tacc, 16);
        if (lane < 16) {
          const int grow = c * BT + row;
          sT[row] = (grow < a.n ? fsys[a.free_idx[grow]] : 0.0) - tacc;
        }
      }
      __syncthreads();

      // ---------------- C: 16x16 diagonal block: L_D L_D^T = P, W = L_D^{-1}  (one warp, in registers)
      if (warp == 0) {
        double* blk = sRing + slot(0, c);
        const int r = lane & 15;
        const int rowpart = ((r >> 3) << 7) + ((r & 7) << 2);
        double row[16];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          const double v = blk[rowpart + ((cc >> 2) << 5) + (cc & 3)];
          row[cc] = lane < 16 ? (cc <= r ? v : 0.0) : (cc == r ? 1.0 : 0.0);
        }
        __syncwarp();
        // lanes 0-15: rows of the block; lanes 16-31: rows of Z = L^{-T} (identity to start with).  The
        // lane owning row k+1 forms the next pivot from its own registers, one shuffle broadcasts it and
        // the rsqrt of column k+1 is issued before the trailing update of column k.
        int bad = 0;
        double d = __shfl_sync(0xffffffffu, row[0], 0);
        if (!(d > 0.0)) bad = c * BT + 1;
        double rinv = rsqrt(bad ? 1.0 : d);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const double lk = row[k] * rinv;                        // lane k: d * rsqrt(d) = sqrt(d)
          row[k] = lk;
          double rinv_next = 0.0;
          if (k < 15) {
            const double dn = __shfl_sync(0xffffffffu, fma(-lk, lk, row[k + 1]), k + 1);
            if (!(dn > 0.0) && !bad) bad = c * BT + k + 2;         // same value in every lane
            rinv_next = rsqrt(bad ? 1.0 : dn);
          }
          double* col = sCol + ((k & 1) << 4);
          if (lane < 16) col[lane] = lk;
          __syncwarp();
#pragma unroll
          for (int cc = k + 1; cc < 16; ++cc) row[cc] = fma(-lk, col[cc], row[cc]);
          rinv = rinv_next;
        }
        if (bad) {
          if (lane == 0) *sFlag = bad;
        } else if (lane >= 16) {
          // W[c'][kk] = Z[kk][c'] for kk <= c' (this lane: kk = r), as a DMMA B operand, over the slot of P
#pragma unroll
          for (int cp = 0; cp < 16; ++cp) blk[b16_off(cp, r)] = (cp >= r) ? row[cp] : 0.0;
        }
      }
      __syncthreads();
      fail = *sFlag;
      if (fail) break;   // uniform

      // ---------------- D: blocks below the diagonal block: L = P W^T (in place + to HBM), W to HBM, y_c
      {
        const double* W = sRing + slot(0, c);
        for (int t = warp; t < 2 * NB; t += TB_BAND_THREADS / 32) {
          const int rb = 1 + (t >> 1), mb = t & 1;
          if (c + rb >= ncol) continue;        // below the last block row: nothing there
          double* blk = sRing + slot(rb, c);
          double a4[4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) a4[ks] = blk[((mb * 4 + ks) << 5) + lane];
          __syncwarp();
          double* g = Lb + ((int64_t)c * (NB + 1) + rb) * BE;
#pragma unroll
          for (int nbp = 0; nbp < 2; ++nbp) {
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma(x0, x1, a4[ks], W[((nbp * 4 + ks) << 5) + lane]);
            const int off = cpair_off(mb, nbp, lane);
            *reinterpret_cast<double2*>(blk + off) = make_double2(x0, x1);
            *reinterpret_cast<double2*>(g + off) = make_double2(x0, x1);
          }
        }
        for (int i = tid; i < BE; i += TB_BAND_THREADS) Wb[(int64_t)c * BE + i] = W[i];
        if (warp == TB_BAND_THREADS / 32 - 1 && lane < 16) {   // y_c = W t  (W lower triangular)
          double yv = 0.0;
          for (int cc = 0; cc <= lane; ++cc) yv = fma(W[b16_off(lane, cc)], sT[cc], yv);
          sY[(c % (NB + 1)) * BT + lane] = yv;
          ysys[c * BT + lane] = yv;
        }
      }
      __syncthreads();
    }

    if (fail) {
      if (tid == 0) a.status[b] = fail;
      __syncthreads();
      continue;
    }

    // ---------------- back substitution: u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb}), last block first
    for (int c = ncol - 1; c >= 0; --c) {
      // stage W_c (slot 0) and the blocks below it (slots 1..NB) from HBM
      for (int i = tid; i < BE; i += TB_BAND_THREADS) sRing[i] = __ldcg(Wb + (int64_t)c * BE + i);
      for (int rb = 1; rb <= NB && c + rb < ncol; ++rb) {
        const double* g = Lb + ((int64_t)c * (NB + 1) + rb) * BE;
        for (int i = tid; i < BE; i += TB_BAND_THREADS) sRing[rb * BE + i] = __ldcg(g + i);
      }
      __syncthreads();
      {
        const int col = tid & 15, part = tid >> 4;   // 8 parts x 2 rows of every block
        double tacc = 0.0;
        for (int rb = 1; rb <= NB && c + rb < ncol; ++rb) {
          const double* blk = sRing + rb * BE;
          const double* uv = sY + ((c + rb) % (NB + 1)) * BT;
          tacc = fma(blk[b16_off(2 * part, col)], uv[2 * part], tacc);
          tacc = fma(blk[b16_off(2 * part + 1, col)], uv[2 * part + 1], tacc);
        }
        sRed[part * BT + col] = tacc;
      }
      __syncthreads();
      if (warp == 0) {
        if (lane < 16) {
          double tsum = 0.0;
#pragma unroll
          for (int p = 0; p < 8; ++p) tsum += sRed[p * BT + lane];
          sT[lane] = __ldcg(ysys + c * BT + lane) - tsum;
        }
        __syncwarp();
        if (lane < 16) {   // u = W^T r: u[col] = sum_{cc >= col} W[cc][col] r[cc]
          double uvv = 0.0;
          for (int cc = lane; cc < 16; ++cc) uvv = fma(sRing[b16_off(cc, lane)], sT[cc], uvv);
          sY[(c % (NB + 1)) * BT + lane] = uvv;
          ysys[c * BT + lane] = uvv;
        }
      }
      __syncthreads();
    }
    if (tid == 0) a.status[b] = 0;
  }
}

}  // namespace

int tb_band_smem_bytes(int NB) {
  const int nring = (NB + 1) * (NB + 2) / 2;
  return (nring * BE + (NB + 1) * BT + BT + 32 + 2) * 8;
}

int tb_launch_band_chol(const LargeArgs& a, int num_sm, cudaStream_t st) {
  const int smem = tb_band_smem_bytes(a.NB);
  cudaError_t e = cudaFuncSetAttribute(k_band, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_band, TB_BAND_THREADS, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sm * per_sm;
  if (grid > a.batch) grid = a.batch;
  tb_prof_begin(TB_PROF_CHOL, st);
  k_band<<<grid, TB_BAND_THREADS, smem, st>>>(a);
  tb_prof_end(TB_PROF_CHOL, st);
  return (int)cudaGetLastError();
}
