// Band path: block-band Cholesky for systems whose reduced stiffness matrix has a narrow
// envelope (bar-942: n = 696, half-bandwidth 56).  The matrix is viewed as a band of 16x16 blocks
// (NB sub-diagonal blocks per block column) and ONE WARP factorises one system, start to finish,
// with no block-level barrier anywhere:
//
//   per block column c:   products   S(R,c)  = sum_{d=1..NB} L(R,c-d) L(c,c-d)^T   (DMMA m8n8k4; the accumulators
//                                     of the whole block column live in registers; operands come from a ring of
//                                     the previous NB block columns in shared memory)
//                         assemble   P(R,c)  = K(R,c) - S(R,c)   (K values prefetched into registers before the
//                                     products, scattered into the ring slots the products just released)
//                         factor     16x16 diagonal block + its inverse W (in registers, lanes = rows)
//                         solve      L(R,c)  = P(R,c) W^T                          (DMMA), kept in the ring + HBM
//                         forward    y_c     = W (f_c - sum_d L(c,c-d) y_{c-d})    (rides on the operand registers)
//   then block back-substitution  u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb}), the factor streamed back
//   from HBM/L2 with cp.async one block column ahead.
//
// Block (p+e, p) is alive from column p to column p+e, so "diagonal e" of the band needs e ring slots
// (slot = e(e-1)/2 + p mod e): NB(NB+1)/2 blocks in all, 20 KB for NB = 4, and ten one-warp CTAs share an SM.
// Structurally zero blocks (block-level symbolic factorisation, one bit mask per block column) are never
// computed, stored or read.  HBM sees the K values once, the non-zero L blocks and the 16x16 inverses
// once out and once back in, and u.  Replaces np.linalg.solve (slientruss3d/truss.py:343) for this class of
// systems; results equal the dense factorisation's (zeros are skipped, nothing else).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "tb_common.cuh"
#include "tb_blocks.cuh"

#ifdef TB_PHASE_TIMING
__device__ unsigned long long g_band_cycles[16];
#define BPH_DECL long long _ph_t = clock64(); unsigned long long _ph_acc[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0};
#define BPH(i) { long long _n = clock64(); _ph_acc[i] += (unsigned long long)(_n - _ph_t); _ph_t = _n; }
#define BPH_FLUSH(cond) if (cond) { for (int _i = 0; _i < 16; ++_i) atomicAdd(&g_band_cycles[_i], _ph_acc[_i]); }
extern "C" int tb_band_phase_read(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, g_band_cycles, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_band_cycles, z, sizeof(z));
  return 0;
}
#else
#define BPH_DECL
#define BPH(i) {}
#define BPH_FLUSH(cond)
#endif

namespace {

constexpr int PRE = 9;    // K entries per lane prefetched into registers per block column (288 per column)
using namespace tbblk;

template <int NB>
struct BandCfg {
  static constexpr int RING = NB * (NB + 1) / 2;                                  // alive off-diagonal blocks
  static constexpr int BUF = RING > 2 * (NB + 1) ? RING : 2 * (NB + 1);           // back substitution: two (NB+1)-block buffers
  static constexpr int DOUBLES = BUF * BE + BE + (NB + 1) * BT + BT + 32 + 8;     // ring | scratch block | y ring | rhs | column | slot table
};

template <int NB, int NW>
__global__ void __launch_bounds__(32 * NW) k_band1(const LargeArgs a) {
  extern __shared__ __align__(16) double sm_all[];
  using Cfg = BandCfg<NB>;
  // NW independent systems per CTA, one per warp, nothing shared between them: the CTA only exists so that the
  // warps are spread over the four schedulers of the SM
  double* sm = sm_all + (threadIdx.x >> 5) * Cfg::DOUBLES;
  double* sRing = sm;                          // blocks of the previous NB block columns | staging of K(:,c)
  double* sScr = sRing + Cfg::BUF * BE;        // diagonal block P(c,c), then W_c
  double* sY = sScr + BE;                      // ring of the last NB+1 blocks of y (then u), 16 each
  double* sT = sY + (NB + 1) * BT;             // [16] right-hand side of the current block
  double* sCol = sT + BT;                      // [32] base case: eliminated column, double buffered
  int* sOff = (int*)(sCol + 32);               // [NB+1] staging slot (in doubles from sm) of block e of this column

  const int lane = threadIdx.x & 31, lsw = lane_swz(lane);
  const int ncol = a.nb16;
  const int qr = lane >> 2, qc = lane & 3;     // this lane's (row, k) inside an operand fragment

  for (int b = blockIdx.x * NW + (threadIdx.x >> 5); b < a.batch; b += gridDim.x * NW) {
    if (a.status[b] != 0) continue;            // input problem flagged by k_geom (uniform per warp)
    const double* kvs = a.kv + (int64_t)b * a.nnz;
    const double* fsys = a.force + b * a.force_stride;
    double* ysys = a.y + (int64_t)b * a.n_pad;
    double* Lb = a.L + (int64_t)b * ncol * (NB + 1) * BE;   // column chunk c: [W_c | L(c+1,c) .. L(c+NB,c)]
    int fail = 0;

    int idx[NB + 1];                           // idx[e] = c mod e
    unsigned nzprev[NB + 1];                   // nzprev[d] = block mask of column c-d
#pragma unroll
    for (int e = 0; e <= NB; ++e) { idx[e] = 0; nzprev[e] = 0u; }
    int yslot = 0;                             // c mod (NB+1)

    // column metadata and K values are fetched one block column ahead, so no load address ever waits on a load
    unsigned nzn = (unsigned)ldg_i32(a.b16_nz);
    int q0n = ldg_i32(a.b16_ptr), q1n = ldg_i32(a.b16_ptr + 2);
    int fin[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) fin[mb] = mb * 8 + qr < a.n ? ldg_i32(a.free_idx + mb * 8 + qr) : -1;
    double kvr[PRE];
    int posr[PRE];
#pragma unroll
    for (int i = 0; i < PRE; ++i) {
      const int q = q0n + lane + 32 * i;
      kvr[i] = 0.0;
      posr[i] = 0;
      if (q < q1n) {
        kvr[i] = ldg_f64(kvs + q);
        posr[i] = ldg_i32(a.b16_pos + q);        // e << 8 | offset inside the block
      }
    }

    for (int c = 0; c < ncol; ++c) {
      const unsigned nzc = nzn;
      const int q0 = q0n, q1 = q1n;
      double fr[2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) fr[mb] = fin[mb] >= 0 ? ldg_f64(fsys + fin[mb]) : 0.0;
      if (c + 1 < ncol) {
        nzn = (unsigned)ldg_i32(a.b16_nz + c + 1);
        q0n = ldg_i32(a.b16_ptr + 2 * (c + 1));
        q1n = ldg_i32(a.b16_ptr + 2 * (c + 2));
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int grow = (c + 1) * BT + mb * 8 + qr;
          fin[mb] = grow < a.n ? ldg_i32(a.free_idx + grow) : -1;
        }
      }

      // ---------------- products with the previous NB block columns (DMMA), forward-substitution partial sums
      double acc[NB + 1][2][2][2];
#pragma unroll
      for (int rb = 0; rb <= NB; ++rb)
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) acc[rb][mb][nb][0] = acc[rb][mb][nb][1] = 0.0;
      // accumulation order of every sum over d: d = 2 .. NB, then d = 1 -- the order in which k_band2's two warps
      // produce the same sums (its trailing warp looks ahead with the terms d >= 2), so both kernels give the same bits
      double tp[2] = {0.0, 0.0}, tp1[2] = {0.0, 0.0};
#pragma unroll
      for (int dq = 0; dq < NB; ++dq) {
        const int d = dq == NB - 1 ? 1 : dq + 2;
        const unsigned nzp = nzprev[d];
        if (!((nzp >> d) & 1u)) continue;      // L(c, c-d) is structurally zero (or c-d < 0): uniform
        const double* Bm = sRing + (d * (d - 1) / 2 + idx[d]) * BE;     // (c-d) mod d == c mod d
        double bf[2][4];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) bf[nb][ks] = Bm[fo(nb * 4 + ks, lsw)];
        {   // rows of L(c, c-d) times y_{c-d}: the operand registers are exactly the needed elements
          int ys = yslot - d;
          if (ys < 0) ys += NB + 1;
          const double* yv = sY + ys * BT;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const double y4 = yv[ks * 4 + qc];
            if (d == 1) {
              tp1[0] = fma(bf[0][ks], y4, tp1[0]);
              tp1[1] = fma(bf[1][ks], y4, tp1[1]);
            } else {
              tp[0] = fma(bf[0][ks], y4, tp[0]);
              tp[1] = fma(bf[1][ks], y4, tp[1]);
            }
          }
        }
#pragma unroll
        for (int rb = 0; rb + d <= NB; ++rb) {
          const int e = rb + d;
          if (!((nzp >> e) & 1u)) continue;    // L(c+rb, c-d) structurally zero: uniform
          int sl = idx[e] - d;                 // (c-d) mod e
          if (sl < 0) sl += e;
          const double* A = sRing + (e * (e - 1) / 2 + sl) * BE;
          double af[2][4];
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) af[mb][ks] = A[fo(mb * 4 + ks, lsw)];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int nb = 0; nb < 2; ++nb)
                if (rb > 0 || nb <= mb) dmma(acc[rb][mb][nb][0], acc[rb][mb][nb][1], af[mb][ks], bf[nb][ks]);
        }
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        tp[mb] += __shfl_xor_sync(0xffffffffu, tp[mb], 1);
        tp[mb] += __shfl_xor_sync(0xffffffffu, tp[mb], 2);
        tp1[mb] += __shfl_xor_sync(0xffffffffu, tp1[mb], 1);
        tp1[mb] += __shfl_xor_sync(0xffffffffu, tp1[mb], 2);
        tp[mb] += tp1[mb];
      }
      __syncwarp();                            // every lane is done with the blocks (c, c-d): their slots are free

      // ---------------- stage K(:,c) into the freed slots (block e -> slot of the dead block (c, c-e); e = 0 -> scratch)
      if (lane == 0) sOff[0] = (int)(sScr - sm);
#pragma unroll
      for (int e = 1; e <= NB; ++e)
        if (lane == e) sOff[e] = (e * (e - 1) / 2 + idx[e]) * BE;
      {
        const double2 z = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(sScr)[lane + 32 * i] = z;
#pragma unroll
        for (int e = 1; e <= NB; ++e)
          if ((nzc >> e) & 1u) {
            double2* blk = reinterpret_cast<double2*>(sRing + (e * (e - 1) / 2 + idx[e]) * BE);
#pragma unroll
            for (int i = 0; i < 4; ++i) blk[lane + 32 * i] = z;
          }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < PRE; ++i)
        if (q0 + lane + 32 * i < q1) sm[sOff[posr[i] >> 8] + (posr[i] & 255)] = kvr[i];
      for (int q = q0 + 32 * PRE + lane; q < q1; q += 32) {   // rare: more than 32*PRE entries in this block column
        const int pos = ldg_i32(a.b16_pos + q);
        sm[sOff[pos >> 8] + (pos & 255)] = ldg_f64(kvs + q);
      }
      if (lane < BT && c * BT + lane >= a.n) sScr[b16_off(lane, lane)] = 1.0;   // identity on the padded diagonal
      if (c + 1 < ncol) {                      // the registers are free again: K values of the next block column
#pragma unroll
        for (int i = 0; i < PRE; ++i) {
          const int q = q0n + lane + 32 * i;
          if (q < q1n) {
            kvr[i] = ldg_f64(kvs + q);
            posr[i] = ldg_i32(a.b16_pos + q);
          }
        }
      }
      __syncwarp();

      // ---------------- P = K - S, in place in the staging slots (fragment layout: the next reads are operand reads)
#pragma unroll
      for (int rb = 0; rb <= NB; ++rb) {
        if (rb > 0 && !((nzc >> rb) & 1u)) continue;
        double* blk = rb == 0 ? sScr : sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            if (rb == 0 && nb > mb) continue;
            double2* p = reinterpret_cast<double2*>(blk + cpair_off(mb, nb, lane));
            double2 v = *p;
            v.x -= acc[rb][mb][nb][0];
            v.y -= acc[rb][mb][nb][1];
            *p = v;
          }
      }
      if (qc == 0) {
        sT[qr] = fr[0] - tp[0];
        sT[8 + qr] = fr[1] - tp[1];
      }
      __syncwarp();

      // ---------------- 16x16 diagonal block: L_D L_D^T = P, W = L_D^{-1}  (in registers)
      fail = factor_diag16(sScr, sCol, lane, c * BT);
      __syncwarp();
      if (fail) break;   // uniform

      // ---------------- y_c = W t;  L(c+rb, c) = P W^T into the ring and to HBM;  W to HBM
      double* chunk = Lb + (int64_t)c * (NB + 1) * BE;
      {
        double yp[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double t4 = sT[ks * 4 + qc];
          yp[0] = fma(sScr[fo(ks, lsw)], t4, yp[0]);
          yp[1] = fma(sScr[fo(4 + ks, lsw)], t4, yp[1]);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          yp[mb] += __shfl_xor_sync(0xffffffffu, yp[mb], 1);
          yp[mb] += __shfl_xor_sync(0xffffffffu, yp[mb], 2);
        }
        if (qc == 0) {
          sY[yslot * BT + qr] = yp[0];
          sY[yslot * BT + 8 + qr] = yp[1];
          ysys[c * BT + qr] = yp[0];
          ysys[c * BT + 8 + qr] = yp[1];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<double2*>(chunk)[lane + 32 * i] = reinterpret_cast<const double2*>(sScr)[lane + 32 * i];
      }
      double wf[2][4];                          // W as the B operand: W[8 nbp + lane/4][4 ks + lane%4]
#pragma unroll
      for (int nbp = 0; nbp < 2; ++nbp)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) wf[nbp][ks] = sScr[fo(nbp * 4 + ks, lsw)];
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nzc >> rb) & 1u)) continue;
        double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE;
        double a4[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) a4[mb][ks] = blk[fo(mb * 4 + ks, lsw)];
        __syncwarp();                          // P(rb) fully read before L(rb) overwrites it
        double* g = chunk + rb * BE;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nbp = 0; nbp < 2; ++nbp) {
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2 * nbp + 2; ++ks) dmma(x0, x1, a4[mb][ks], wf[nbp][ks]);   // W is lower triangular
            const int off = cpair_off(mb, nbp, lane);
            *reinterpret_cast<double2*>(blk + off) = make_double2(x0, x1);
            *reinterpret_cast<double2*>(g + off) = make_double2(x0, x1);
          }
      }
      __syncwarp();

      // ---------------- advance the ring counters
#pragma unroll
      for (int e = NB; e >= 2; --e) nzprev[e] = nzprev[e - 1];
      nzprev[1] = nzc;
#pragma unroll
      for (int e = 1; e <= NB; ++e) idx[e] = (idx[e] + 1 == e) ? 0 : idx[e] + 1;
      yslot = (yslot == NB) ? 0 : yslot + 1;
    }

    if (fail) {
      if (lane == 0) a.status[b] = fail;
      __syncwarp();
      continue;
    }

    // ---------------- back substitution: u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb}), last block first.
    // Column chunks come back from HBM/L2 through a two-buffer cp.async pipeline in the (now idle) ring.
    auto fetch = [&](int c) {
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c) | 1u;     // bit 0: W_c
      const double* src = Lb + (int64_t)c * (NB + 1) * BE;
      double* dst = sRing + (c & 1) * (NB + 1) * BE;
#pragma unroll
      for (int e = 0; e <= NB; ++e)
        if ((nz >> e) & 1u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async16(dst + e * BE + (lane + 32 * i) * 2, src + e * BE + (lane + 32 * i) * 2);
        }
      cp_async_commit();
    };
    __syncwarp();
    fetch(ncol - 1);
    // y_c and the block mask of column c are fetched one block column ahead (they come from L2)
    double ycn[4] = {0.0, 0.0, 0.0, 0.0};
    if (qr == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (ncol - 1) * BT + ks * 4 + qc);
    }
    unsigned nzb = (unsigned)ldg_i32(a.b16_nz + ncol - 1);
    // after the factorisation yslot == ncol mod (NB+1); block column c of y sits in slot c mod (NB+1)
    for (int c = ncol - 1; c >= 0; --c) {
      yslot = (yslot == 0) ? NB : yslot - 1;   // slot of block c (the ring only keeps the last NB+1 blocks: y_c comes from HBM)
      double yc[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) yc[ks] = ycn[ks];
      const unsigned nz = nzb;
      if (c > 0) {
        nzb = (unsigned)ldg_i32(a.b16_nz + c - 1);
        if (qr == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (c - 1) * BT + ks * 4 + qc);
        }
      }
      if (c > 0) {
        fetch(c - 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const double* buf = sRing + (c & 1) * (NB + 1) * BE;
      // t[col] = sum_rb sum_r L(c+rb,c)[r][col] u_{c+rb}[r]; this lane: r = 8 mb + lane/4, col = 4 ks + lane%4
      double tp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nz >> rb) & 1u)) continue;
        int us = yslot + rb;
        if (us > NB) us -= NB + 1;
        const double* uv = sY + us * BT;
        const double* blk = buf + rb * BE;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const double ur = uv[mb * 8 + qr];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tp[ks] = fma(blk[fo(mb * 4 + ks, lsw)], ur, tp[ks]);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 4);
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 8);
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 16);
      }
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) sT[ks * 4 + qc] = yc[ks] - tp[ks];
      }
      __syncwarp();
      // u = W^T r: u[col] = sum_{cc >= col} W[cc][col] r[cc]
      double up[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const double rr = sT[mb * 8 + qr];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) up[ks] = fma(buf[fo(mb * 4 + ks, lsw)], rr, up[ks]);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 4);
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 8);
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 16);
      }
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          sY[yslot * BT + ks * 4 + qc] = up[ks];
          ysys[c * BT + ks * 4 + qc] = up[ks];
        }
      }
      __syncwarp();
    }
    if (lane == 0) a.status[b] = 0;
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------------------------
// Two warps per system (one CTA = one system), for batches too small to fill the GPU with one warp per system
// (1024 systems on 148 SMs).  Same algorithm, data layout and arithmetic as k_band1; the block column is split
// into the latency chain and the throughput work, which then overlap:
//
//   warp F  diagonal products S(c,c) -> P(c,c) -> 16x16 factorisation + W_c -> y_c        | X | L(c+1,c) = P W^T | Y |
//   warp T  off-diagonal products S(c+rb,c) (96 of the 144 DMMAs) -> stage K -> P(c+rb,c) | X | L(c+rb,c), rb>=2 | Y |
//
// X and Y are CTA barriers (W_c and the P blocks exist / column c of L is in the ring); one more arrive/sync pair
// (Z) tells T that F no longer reads the blocks (c, c-d) whose slots take the staged K blocks.  The back
// substitution runs on warp F alone.  No instruction is duplicated between the warps except the operand loads of
// L(c, c-d), so the instruction count per system stays that of k_band1 while the pivot chain of column c hides the
// tensor work of the same column.
// ------------------------------------------------------------------------------------------------------------
constexpr int PREF = 5;   // diagonal-block K entries per lane held in registers (<= 136 entries)
constexpr int PRET = 6;   // off-diagonal K entries per lane held in registers

__device__ __forceinline__ void bar_arrive_z() { asm volatile("bar.arrive 1, 64;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_z() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

template <int NB>
__global__ void __launch_bounds__(64, 7) k_band2(const LargeArgs a) {
  extern __shared__ __align__(16) double sm[];
  using Cfg = BandCfg<NB>;
  double* sRing = sm;
  double* sScr = sRing + Cfg::BUF * BE;
  double* sY = sScr + BE;
  double* sT = sY + (NB + 1) * BT;
  double* sCol = sT + BT;
  int* sOff = (int*)(sCol + 32);               // [NB+1] staging slots, [NB+1] failure flag
  double* sSd = sCol + 32 + 8;                 // [3][32][2] warp T's part of S(c+1,c+1) (terms d >= 2), accumulator layout
  double* sTp = sSd + 192;                     // [16] warp T's part of sum_d L(c+1,c+1-d) y_{c+1-d} (terms d >= 2)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lsw = lane_swz(lane);
  const int ncol = a.nb16;
  const int qr = lane >> 2, qc = lane & 3;

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (a.status[b] != 0) continue;            // input problem flagged by k_geom (uniform)
    const bool isF = warp == (b & 1);          // alternate the roles so co-resident CTAs load the schedulers evenly
    const double* kvs = a.kv + (int64_t)b * a.nnz;
    const double* fsys = a.force + b * a.force_stride;
    double* ysys = a.y + (int64_t)b * a.n_pad;
    double* Lb = a.L + (int64_t)b * ncol * (NB + 1) * BE;
    int fail = 0;
    if (tid == 0) sOff[NB + 1] = 0;

    int idx[NB + 1];
    unsigned nzprev[NB + 1];
#pragma unroll
    for (int e = 0; e <= NB; ++e) { idx[e] = 0; nzprev[e] = 0u; }
    int yslot = 0;
    // column metadata is fetched one block column ahead, so that no load address waits on another load
    const int part = isF ? 0 : 1;              // F: entries of the diagonal block; T: entries below it
    unsigned nzn = (unsigned)ldg_i32(a.b16_nz);
    int q0n = ldg_i32(a.b16_ptr + part), q1n = ldg_i32(a.b16_ptr + part + 1);
    int fin[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) fin[h] = (isF && h * 8 + qr < a.n) ? ldg_i32(a.free_idx + h * 8 + qr) : -1;
    BPH_DECL
    __syncthreads();

    for (int c = 0; c < ncol; ++c) {
      const unsigned nzc = nzn;
      const int q0 = q0n, q1 = q1n;
      const int fi0 = fin[0], fi1 = fin[1];
      if (c + 1 < ncol) {
        nzn = (unsigned)ldg_i32(a.b16_nz + c + 1);
        q0n = ldg_i32(a.b16_ptr + 2 * (c + 1) + part);
        q1n = ldg_i32(a.b16_ptr + 2 * (c + 1) + part + 1);
        if (isF) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int grow = (c + 1) * BT + h * 8 + qr;
            fin[h] = grow < a.n ? ldg_i32(a.free_idx + grow) : -1;
          }
        }
      }
      double* chunk = Lb + (int64_t)c * (NB + 1) * BE;

      if (isF) {
        // ================= warp F: the latency chain of block column c =================
        double kvf[PREF];
        int posf[PREF];
#pragma unroll
        for (int i = 0; i < PREF; ++i) {
          const int q = q0 + lane + 32 * i;
          kvf[i] = 0.0;
          posf[i] = 0;
          if (q < q1) {
            kvf[i] = ldg_f64(kvs + q);
            posf[i] = ldg_i32(a.b16_pos + q);
          }
        }
        double fr[2];
        fr[0] = fi0 >= 0 ? ldg_f64(fsys + fi0) : 0.0;
        fr[1] = fi1 >= 0 ? ldg_f64(fsys + fi1) : 0.0;
        // sub-blocks (0,0), (1,0), (1,1) of S(c,c) and the forward-substitution sums: warp T left the terms d >= 2 in
        // shared memory during the previous block column (they only need columns <= c-2); the d = 1 term needs
        // L(c,c-1), which this warp produced at the end of the previous block column
        double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        double tp[2] = {0.0, 0.0};
        if (c > 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double2 v = reinterpret_cast<const double2*>(sSd)[q * 32 + lane];
            acc[q][0] = v.x;
            acc[q][1] = v.y;
          }
          tp[0] = sTp[qr];
          tp[1] = sTp[8 + qr];
        }
        if ((nzprev[1] >> 1) & 1u) {
          const double* Bm = sRing + idx[1] * BE;       // slot of block (c, c-1): diagonal 1 has one slot
          double bf[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) bf[h][ks] = Bm[fo(h * 4 + ks, lsw)];
          int ys = yslot - 1;
          if (ys < 0) ys += NB + 1;
          const double* yv = sY + ys * BT;
          double t0 = 0.0, t1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const double y4 = yv[ks * 4 + qc];
            t0 = fma(bf[0][ks], y4, t0);
            t1 = fma(bf[1][ks], y4, t1);
          }
          t0 += __shfl_xor_sync(0xffffffffu, t0, 1);
          t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
          tp[0] += t0;
          tp[1] += t1;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // L(c,c-1) L(c,c-1)^T: the operand fragments of A and B coincide
            dmma(acc[0][0], acc[0][1], bf[0][ks], bf[0][ks]);
            dmma(acc[1][0], acc[1][1], bf[1][ks], bf[0][ks]);
            dmma(acc[2][0], acc[2][1], bf[1][ks], bf[1][ks]);
          }
        }
        __syncwarp();
        bar_arrive_z();                        // [Z] F no longer reads the blocks (c, c-d)
        BPH(0)
        {
          const double2 z = make_double2(0.0, 0.0);
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(sScr)[lane + 32 * i] = z;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PREF; ++i)
          if (q0 + lane + 32 * i < q1) sScr[posf[i] & 255] = kvf[i];
        for (int q = q0 + 32 * PREF + lane; q < q1; q += 32) sScr[ldg_i32(a.b16_pos + q) & 255] = ldg_f64(kvs + q);
        if (lane < BT && c * BT + lane >= a.n) sScr[b16_off(lane, lane)] = 1.0;   // identity on the padded diagonal
        __syncwarp();
        {
          double2* p0 = reinterpret_cast<double2*>(sScr + cpair_off(0, 0, lane));
          double2* p1 = reinterpret_cast<double2*>(sScr + cpair_off(1, 0, lane));
          double2* p2 = reinterpret_cast<double2*>(sScr + cpair_off(1, 1, lane));
          double2 v0 = *p0, v1 = *p1, v2 = *p2;
          v0.x -= acc[0][0]; v0.y -= acc[0][1];
          v1.x -= acc[1][0]; v1.y -= acc[1][1];
          v2.x -= acc[2][0]; v2.y -= acc[2][1];
          *p0 = v0; *p1 = v1; *p2 = v2;
        }
        if (qc == 0) {
          sT[qr] = fr[0] - tp[0];
          sT[8 + qr] = fr[1] - tp[1];
        }
        __syncwarp();
        BPH(1)
        const int bad = factor_diag16(sScr, sCol, lane, c * BT);
        if (bad && lane == 0) sOff[NB + 1] = bad;
        __syncwarp();
        BPH(2)
        if (!bad) {                            // y_c = W t, W_c to HBM
          double yp[2] = {0.0, 0.0};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const double t4 = sT[ks * 4 + qc];
            yp[0] = fma(sScr[fo(ks, lsw)], t4, yp[0]);
            yp[1] = fma(sScr[fo(4 + ks, lsw)], t4, yp[1]);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 1);
            yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 2);
          }
          if (qc == 0) {
            sY[yslot * BT + qr] = yp[0];
            sY[yslot * BT + 8 + qr] = yp[1];
            ysys[c * BT + qr] = yp[0];
            ysys[c * BT + 8 + qr] = yp[1];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<double2*>(chunk)[lane + 32 * i] = reinterpret_cast<const double2*>(sScr)[lane + 32 * i];
        }
        BPH(3)
      } else {
        // ================= warp T: the tensor work of block column c =================
        double kvr[PRET];
        int posr[PRET];
#pragma unroll
        for (int i = 0; i < PRET; ++i) {
          const int q = q0 + lane + 32 * i;
          kvr[i] = 0.0;
          posr[i] = 0;
          if (q < q1) {
            kvr[i] = ldg_f64(kvs + q);
            posr[i] = ldg_i32(a.b16_pos + q);
          }
        }
#pragma unroll
        for (int e = 1; e <= NB; ++e)
          if (lane == e) sOff[e] = (e * (e - 1) / 2 + idx[e]) * BE;
        double acc[NB + 1][2][2][2];
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) acc[rb][mb][nb][0] = acc[rb][mb][nb][1] = 0.0;
#pragma unroll
        for (int dq = 0; dq < NB - 1; ++dq) {    // d = 2 .. NB-1, then d = 1 (k_band1's order: same bits from both kernels)
          const int d = dq == NB - 2 ? 1 : dq + 2;
          const unsigned nzp = nzprev[d];
          if (!((nzp >> d) & 1u) || !(nzp >> (d + 1))) continue;   // L(c,c-d) zero, or nothing below it in that column
          const double* Bm = sRing + (d * (d - 1) / 2 + idx[d]) * BE;
          double bf[2][4];
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) bf[nb][ks] = Bm[fo(nb * 4 + ks, lsw)];
#pragma unroll
          for (int rb = 1; rb + d <= NB; ++rb) {
            const int e = rb + d;
            if (!((nzp >> e) & 1u)) continue;
            int sl = idx[e] - d;               // (c-d) mod e
            if (sl < 0) sl += e;
            const double* A = sRing + (e * (e - 1) / 2 + sl) * BE;
            double af[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) af[mb][ks] = A[fo(mb * 4 + ks, lsw)];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
              for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) dmma(acc[rb][mb][nb][0], acc[rb][mb][nb][1], af[mb][ks], bf[nb][ks]);
          }
        }
        __syncwarp();
        bar_sync_z();                          // [Z] F is done with the blocks (c, c-d): their slots take K(c+e, c)
        {
          const double2 z = make_double2(0.0, 0.0);
#pragma unroll
          for (int e = 1; e <= NB; ++e)
            if ((nzc >> e) & 1u) {
              double2* blk = reinterpret_cast<double2*>(sRing + (e * (e - 1) / 2 + idx[e]) * BE);
#pragma unroll
              for (int i = 0; i < 4; ++i) blk[lane + 32 * i] = z;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PRET; ++i)
          if (q0 + lane + 32 * i < q1) sm[sOff[posr[i] >> 8] + (posr[i] & 255)] = kvr[i];
        for (int q = q0 + 32 * PRET + lane; q < q1; q += 32) {
          const int pos = ldg_i32(a.b16_pos + q);
          sm[sOff[pos >> 8] + (pos & 255)] = ldg_f64(kvs + q);
        }
        __syncwarp();
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb) {
          if (!((nzc >> rb) & 1u)) continue;
          double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE;
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              double2* p = reinterpret_cast<double2*>(blk + cpair_off(mb, nb, lane));
              double2 v = *p;
              v.x -= acc[rb][mb][nb][0];
              v.y -= acc[rb][mb][nb][1];
              *p = v;
            }
        }
        // ---- look-ahead for warp F: terms d >= 2 of S(c+1,c+1) and of the forward-substitution sum of block column
        // c+1.  Block (c+1, c+1-d) is block (c + rb, c - dd) with rb = 1, dd = d - 1: slot and mask as seen from column c.
        {
          double sd[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
          double tq[2] = {0.0, 0.0};
#pragma unroll
          for (int dd = 1; dd < NB; ++dd) {
            const int e = dd + 1;
            if (!((nzprev[dd] >> e) & 1u)) continue;
            int sl = idx[e] - dd;              // (c-dd) mod e
            if (sl < 0) sl += e;
            const double* Bm = sRing + (e * (e - 1) / 2 + sl) * BE;
            double bf[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) bf[h][ks] = Bm[fo(h * 4 + ks, lsw)];
            int ys = yslot - dd;               // y_{c-dd}
            if (ys < 0) ys += NB + 1;
            const double* yv = sY + ys * BT;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const double y4 = yv[ks * 4 + qc];
              tq[0] = fma(bf[0][ks], y4, tq[0]);
              tq[1] = fma(bf[1][ks], y4, tq[1]);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              dmma(sd[0][0], sd[0][1], bf[0][ks], bf[0][ks]);
              dmma(sd[1][0], sd[1][1], bf[1][ks], bf[0][ks]);
              dmma(sd[2][0], sd[2][1], bf[1][ks], bf[1][ks]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tq[h] += __shfl_xor_sync(0xffffffffu, tq[h], 1);
            tq[h] += __shfl_xor_sync(0xffffffffu, tq[h], 2);
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) reinterpret_cast<double2*>(sSd)[q * 32 + lane] = make_double2(sd[q][0], sd[q][1]);
          if (qc == 0) {
            sTp[qr] = tq[0];
            sTp[8 + qr] = tq[1];
          }
        }
      }
      __syncthreads();                         // [X] W_c is in the scratch block, P(c+rb, c) are in their slots
      BPH(4)
      fail = sOff[NB + 1];
      if (fail) break;   // uniform

      // ---------------- L(c+rb, c) = P W^T into the ring and to HBM: warp F takes rb = 1, warp T the rest
      {
        double wf[2][4];
#pragma unroll
        for (int nbp = 0; nbp < 2; ++nbp)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) wf[nbp][ks] = sScr[fo(nbp * 4 + ks, lsw)];
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb) {
          if ((rb == 1) != isF) continue;
          if (!((nzc >> rb) & 1u)) continue;
          double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE;
          double a4[2][4];
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) a4[mb][ks] = blk[fo(mb * 4 + ks, lsw)];
          __syncwarp();                        // P(rb) fully read before L(rb) overwrites it
          double* g = chunk + rb * BE;
          double x[2][2][2] = {};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int nbp = 0; nbp < 2; ++nbp)
                if (ks < 2 * nbp + 2) dmma(x[mb][nbp][0], x[mb][nbp][1], a4[mb][ks], wf[nbp][ks]);   // W is lower triangular
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nbp = 0; nbp < 2; ++nbp) {
              const int off = cpair_off(mb, nbp, lane);
              *reinterpret_cast<double2*>(blk + off) = make_double2(x[mb][nbp][0], x[mb][nbp][1]);
              *reinterpret_cast<double2*>(g + off) = make_double2(x[mb][nbp][0], x[mb][nbp][1]);
            }
        }
      }
      BPH(5)
      __syncthreads();                         // [Y] column c of L is in the ring
      BPH(6)

#pragma unroll
      for (int e = NB; e >= 2; --e) nzprev[e] = nzprev[e - 1];
      nzprev[1] = nzc;
#pragma unroll
      for (int e = 1; e <= NB; ++e) idx[e] = (idx[e] + 1 == e) ? 0 : idx[e] + 1;
      yslot = (yslot == NB) ? 0 : yslot + 1;
    }

    if (fail) {
      if (tid == 0) a.status[b] = fail;
      __syncthreads();
      continue;
    }

    // ---------------- back substitution on warp F: u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb}), last block first
    if (isF) {
      auto fetch = [&](int c) {
        const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c) | 1u;     // bit 0: W_c
        const double* src = Lb + (int64_t)c * (NB + 1) * BE;
        double* dst = sRing + (c & 1) * (NB + 1) * BE;
#pragma unroll
        for (int e = 0; e <= NB; ++e)
          if ((nz >> e) & 1u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async16(dst + e * BE + (lane + 32 * i) * 2, src + e * BE + (lane + 32 * i) * 2);
          }
        cp_async_commit();
      };
      fetch(ncol - 1);
      // y_c and the block mask of column c are fetched one block column ahead (they come from L2)
      double ycn[4] = {0.0, 0.0, 0.0, 0.0};
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (ncol - 1) * BT + ks * 4 + qc);
      }
      unsigned nzb = (unsigned)ldg_i32(a.b16_nz + ncol - 1);
      for (int c = ncol - 1; c >= 0; --c) {
        yslot = (yslot == 0) ? NB : yslot - 1;
        double yc[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) yc[ks] = ycn[ks];
        const unsigned nz = nzb;
        if (c > 0) {
          nzb = (unsigned)ldg_i32(a.b16_nz + c - 1);
          if (qr == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (c - 1) * BT + ks * 4 + qc);
          }
        }
        if (c > 0) {
          fetch(c - 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const double* buf = sRing + (c & 1) * (NB + 1) * BE;
        double tp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb) {
          if (!((nz >> rb) & 1u)) continue;
          int us = yslot + rb;
          if (us > NB) us -= NB + 1;
          const double* uv = sY + us * BT;
          const double* blk = buf + rb * BE;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double ur = uv[h * 8 + qr];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tp[ks] = fma(blk[fo(h * 4 + ks, lsw)], ur, tp[ks]);
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 4);
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 8);
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 16);
        }
        if (qr == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) sT[ks * 4 + qc] = yc[ks] - tp[ks];
        }
        __syncwarp();
        double up[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double rr = sT[h * 8 + qr];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) up[ks] = fma(buf[fo(h * 4 + ks, lsw)], rr, up[ks]);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 4);
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 8);
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 16);
        }
        if (qr == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            sY[yslot * BT + ks * 4 + qc] = up[ks];
            ysys[c * BT + ks * 4 + qc] = up[ks];
          }
        }
        __syncwarp();
      }
      BPH(7)
      if (lane == 0) a.status[b] = 0;
    }
    BPH_FLUSH(lane == 0 && isF)
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Three warps per system.  Same arithmetic, operand layout and summation orders as k_band1 / k_band2 (the three kernels
// return the same bits); the block column is cut so that warp F carries the pivot chain and nothing else:
//
//   warp F  d = 1 term of S(c,c) -> P(c,c) = K - S -> 16x16 factorisation + W_c          | X | L(c+1,c) = P W^T     | Y |
//   warp T  S(c+rb,c) -> [Z] -> -S into the staging slots -> [G] -> look-ahead S(c+1,c+1) | X | L(c+rb,c), rb >= 2   | Y |
//   warp G  K values (global -> registers), column c-1 of L and W_{c-1} -> HBM, forward-substitution sums -> [Z]
//           -> K(c+1,c+1) into the other scratch block -> [G] -> P(c+rb,c) = -S + K       | X | y_c = W_c t_c        | Y |
//
// X, Y: CTA barriers.  Z (named barrier 1): F and G no longer read the blocks (c, c-d) whose slots take column c.
// G (named barrier 2): -S is in the staging slots.  Nothing from HBM, no zero fill and no scatter sits on warp F's
// path any more; warp T no longer holds K values in registers; every global store is issued by warp G from shared
// memory one block column later.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NB>
struct Band3Cfg {
  static constexpr int BUF = BandCfg<NB>::BUF;
  // ring | two scratch blocks | y ring | rhs | column | slot table, failure flag | look-ahead part of S(c+1,c+1)
  static constexpr int DOUBLES = BUF * BE + 2 * BE + (NB + 1) * BT + BT + 32 + 8 + 192;
};

#ifdef TB_BAND3_80   // experiment: cap at 80 registers so that seven (eight) systems fit an SM
#define TB_BAND3_ATTR __launch_bounds__(96, 7)
#else
#define TB_BAND3_ATTR __maxnreg__(96)
#endif
template <int NB>
__global__ void TB_BAND3_ATTR k_band3(const LargeArgs a) {
  extern __shared__ __align__(16) double sm[];
  using Cfg = Band3Cfg<NB>;
  double* sRing = sm;
  double* sScr0 = sRing + Cfg::BUF * BE;       // block column c uses scratch block (c & 1): P(c,c), then W_c
  double* sY = sScr0 + 2 * BE;
  double* sT = sY + (NB + 1) * BT;
  double* sCol = sT + BT;
  int* sOff = (int*)(sCol + 32);               // [1..NB] staging slots, [NB+1] failure flag
  double* sSd = sCol + 32 + 8;                 // [3][32][2] terms d >= 2 of S(c+1,c+1), accumulator layout

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lsw = lane_swz(lane);
  const int ncol = a.nb16;
  const int qr = lane >> 2, qc = lane & 3;

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (a.status[b] != 0) continue;            // input problem flagged by k_geom (uniform)
    int role = warp + (b % 3);                 // rotate the roles so co-resident CTAs load the schedulers evenly
    if (role >= 3) role -= 3;
    const bool isF = role == 0, isT = role == 1, isG = role == 2;
    const double* kvs = a.kv + (int64_t)b * a.nnz;
    const double* fsys = a.force + b * a.force_stride;
    double* ysys = a.y + (int64_t)b * a.n_pad;
    double* Lb = a.L + (int64_t)b * ncol * (NB + 1) * BE;
    int fail = 0;
    if (tid == 0) sOff[NB + 1] = 0;

    int idx[NB + 1];
    unsigned nzprev[NB + 1];
#pragma unroll
    for (int e = 0; e <= NB; ++e) { idx[e] = 0; nzprev[e] = 0u; }
    int yslot = 0;
    unsigned nzn = (unsigned)ldg_i32(a.b16_nz);
    // warp G: entry ranges [po0, po1) below the diagonal block of column c, [po1, pd1) diagonal block of column c+1
    int po0 = 0, po1 = 0, pd1 = 0, pn1 = 0, pn2 = 0;
    int fin[2] = {-1, -1};
    if (isG) {
      const int d0 = ldg_i32(a.b16_ptr);
      po0 = ldg_i32(a.b16_ptr + 1);
      po1 = ldg_i32(a.b16_ptr + 2);
      pd1 = ncol > 1 ? ldg_i32(a.b16_ptr + 3) : po1;
#pragma unroll
      for (int h = 0; h < 2; ++h) fin[h] = h * 8 + qr < a.n ? ldg_i32(a.free_idx + h * 8 + qr) : -1;
      // K(0,0) into scratch block 0
      const double2 z = make_double2(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(sScr0)[lane + 32 * i] = z;
      __syncwarp();
      for (int q = d0 + lane; q < po0; q += 32) sScr0[ldg_i32(a.b16_pos + q) & 255] = ldg_f64(kvs + q);
      if (lane < BT && lane >= a.n) sScr0[b16_off(lane, lane)] = 1.0;
    }
    BPH_DECL
    __syncthreads();

    for (int c = 0; c < ncol; ++c) {
      const unsigned nzc = nzn;
      if (c + 1 < ncol) nzn = (unsigned)ldg_i32(a.b16_nz + c + 1);
      double* scr = sScr0 + (c & 1) * BE;

      if (isF) {
        // ================= warp F: the pivot chain of block column c =================
        double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        if (c > 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double2 v = reinterpret_cast<const double2*>(sSd)[q * 32 + lane];
            acc[q][0] = v.x;
            acc[q][1] = v.y;
          }
        }
        if ((nzprev[1] >> 1) & 1u) {
          const double* Bm = sRing + idx[1] * BE;       // block (c, c-1): diagonal 1 has one slot
          double bf[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) bf[h][ks] = Bm[fo(h * 4 + ks, lsw)];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // L(c,c-1) L(c,c-1)^T: the operand fragments of A and B coincide
            dmma(acc[0][0], acc[0][1], bf[0][ks], bf[0][ks]);
            dmma(acc[1][0], acc[1][1], bf[1][ks], bf[0][ks]);
            dmma(acc[2][0], acc[2][1], bf[1][ks], bf[1][ks]);
          }
        }
        __syncwarp();
        bar_arrive_n(1, 96);                   // [Z] F no longer reads block (c, c-1)
        BPH(0)
        {
          double2* p0 = reinterpret_cast<double2*>(scr + cpair_off(0, 0, lane));
          double2* p1 = reinterpret_cast<double2*>(scr + cpair_off(1, 0, lane));
          double2* p2 = reinterpret_cast<double2*>(scr + cpair_off(1, 1, lane));
          double2 v0 = *p0, v1 = *p1, v2 = *p2;
          v0.x -= acc[0][0]; v0.y -= acc[0][1];
          v1.x -= acc[1][0]; v1.y -= acc[1][1];
          v2.x -= acc[2][0]; v2.y -= acc[2][1];
          *p0 = v0; *p1 = v1; *p2 = v2;
        }
        __syncwarp();
        BPH(1)
        const int bad = factor_diag16(scr, sCol, lane, c * BT);
        if (bad && lane == 0) sOff[NB + 1] = bad;
        __syncwarp();
        BPH(2)
      } else if (isT) {
        // ================= warp T: the tensor work of block column c =================
        double acc[NB + 1][2][2][2];
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) acc[rb][mb][nb][0] = acc[rb][mb][nb][1] = 0.0;
#pragma unroll
        for (int dq = 0; dq < NB - 1; ++dq) {    // d = 2 .. NB-1, then d = 1 (k_band1's order)
          const int d = dq == NB - 2 ? 1 : dq + 2;
          const unsigned nzp = nzprev[d];
          if (!((nzp >> d) & 1u) || !(nzp >> (d + 1))) continue;   // L(c,c-d) zero, or nothing below it in that column
          const double* Bm = sRing + (d * (d - 1) / 2 + idx[d]) * BE;
          double bf[2][4];
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) bf[nb][ks] = Bm[fo(nb * 4 + ks, lsw)];
#pragma unroll
          for (int rb = 1; rb + d <= NB; ++rb) {
            const int e = rb + d;
            if (!((nzp >> e) & 1u)) continue;
            int sl = idx[e] - d;               // (c-d) mod e
            if (sl < 0) sl += e;
            const double* A = sRing + (e * (e - 1) / 2 + sl) * BE;
            double af[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) af[mb][ks] = A[fo(mb * 4 + ks, lsw)];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
              for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) dmma(acc[rb][mb][nb][0], acc[rb][mb][nb][1], af[mb][ks], bf[nb][ks]);
          }
        }
        __syncwarp();
        bar_sync_n(1, 96);                     // [Z] the blocks (c, c-d) are dead: their slots take column c
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb) {
          if (!((nzc >> rb) & 1u)) continue;
          double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE;
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
              *reinterpret_cast<double2*>(blk + cpair_off(mb, nb, lane)) =
                  make_double2(0.0 - acc[rb][mb][nb][0], 0.0 - acc[rb][mb][nb][1]);
        }
        __syncwarp();
        bar_arrive_n(2, 64);                   // [G] -S is in the staging slots
        // ---- look-ahead for warp F: terms d >= 2 of S(c+1,c+1).  Block (c+1, c+1-d) is block (c + 1, c - dd), dd = d - 1
        {
          double sd[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
          for (int dd = 1; dd < NB; ++dd) {
            const int e = dd + 1;
            if (!((nzprev[dd] >> e) & 1u)) continue;
            int sl = idx[e] - dd;              // (c-dd) mod e
            if (sl < 0) sl += e;
            const double* Bm = sRing + (e * (e - 1) / 2 + sl) * BE;
            double bf[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) bf[h][ks] = Bm[fo(h * 4 + ks, lsw)];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              dmma(sd[0][0], sd[0][1], bf[0][ks], bf[0][ks]);
              dmma(sd[1][0], sd[1][1], bf[1][ks], bf[0][ks]);
              dmma(sd[2][0], sd[2][1], bf[1][ks], bf[1][ks]);
            }
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) reinterpret_cast<double2*>(sSd)[q * 32 + lane] = make_double2(sd[q][0], sd[q][1]);
        }
      } else {
        // ================= warp G: K values, forward substitution, stores =================
        double kvo[PRET], kvd[PREF];
        int poso[PRET], posd[PREF];
#pragma unroll
        for (int i = 0; i < PRET; ++i) {
          const int q = po0 + lane + 32 * i;
          kvo[i] = 0.0;
          poso[i] = 0;
          if (q < po1) {
            kvo[i] = ldg_f64(kvs + q);
            poso[i] = ldg_i32(a.b16_pos + q);
          }
        }
#pragma unroll
        for (int i = 0; i < PREF; ++i) {
          const int q = po1 + lane + 32 * i;
          kvd[i] = 0.0;
          posd[i] = 0;
          if (q < pd1) {
            kvd[i] = ldg_f64(kvs + q);
            posd[i] = ldg_i32(a.b16_pos + q);
          }
        }
        double fr[2];
        fr[0] = fin[0] >= 0 ? ldg_f64(fsys + fin[0]) : 0.0;
        fr[1] = fin[1] >= 0 ? ldg_f64(fsys + fin[1]) : 0.0;
        if (c + 1 < ncol) {                    // metadata of the next block column
          pn1 = ldg_i32(a.b16_ptr + 2 * c + 4);
          pn2 = c + 2 < ncol ? ldg_i32(a.b16_ptr + 2 * c + 5) : -1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int grow = (c + 1) * BT + h * 8 + qr;
            fin[h] = grow < a.n ? ldg_i32(a.free_idx + grow) : -1;
          }
        }
#pragma unroll
        for (int e = 1; e <= NB; ++e)
          if (lane == e) sOff[e] = (e * (e - 1) / 2 + idx[e]) * BE;
        // ---- block column c-1 of the factor -> HBM (its blocks are in the ring, W_{c-1} in the other scratch block)
        if (c > 0) {
          double* chunk = Lb + (int64_t)(c - 1) * (NB + 1) * BE;
          const double* wsrc = sScr0 + ((c - 1) & 1) * BE;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<double2*>(chunk)[lane + 32 * i] = reinterpret_cast<const double2*>(wsrc)[lane + 32 * i];
#pragma unroll
          for (int rb = 1; rb <= NB; ++rb) {
            if (!((nzprev[1] >> rb) & 1u)) continue;
            const int sl = idx[rb] == 0 ? rb - 1 : idx[rb] - 1;          // (c-1) mod rb
            const double* src = sRing + (rb * (rb - 1) / 2 + sl) * BE;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<double2*>(chunk + rb * BE)[lane + 32 * i] = reinterpret_cast<const double2*>(src)[lane + 32 * i];
          }
        }
        // ---- forward substitution: sum_d L(c,c-d) y_{c-d}, terms d = 2..NB then d = 1 (k_band1's order)
        double tp[2] = {0.0, 0.0}, tp1[2] = {0.0, 0.0};
#pragma unroll
        for (int dq = 0; dq < NB; ++dq) {
          const int d = dq == NB - 1 ? 1 : dq + 2;
          if (!((nzprev[d] >> d) & 1u)) continue;
          const double* Bm = sRing + (d * (d - 1) / 2 + idx[d]) * BE;
          int ys = yslot - d;
          if (ys < 0) ys += NB + 1;
          const double* yv = sY + ys * BT;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const double y4 = yv[ks * 4 + qc];
            const double b0 = Bm[fo(ks, lsw)], b1 = Bm[fo(4 + ks, lsw)];
            if (d == 1) {
              tp1[0] = fma(b0, y4, tp1[0]);
              tp1[1] = fma(b1, y4, tp1[1]);
            } else {
              tp[0] = fma(b0, y4, tp[0]);
              tp[1] = fma(b1, y4, tp[1]);
            }
          }
        }
        __syncwarp();
        bar_arrive_n(1, 96);                   // [Z] G no longer reads the blocks (c, c-d)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tp[h] += __shfl_xor_sync(0xffffffffu, tp[h], 1);
          tp[h] += __shfl_xor_sync(0xffffffffu, tp[h], 2);
          tp1[h] += __shfl_xor_sync(0xffffffffu, tp1[h], 1);
          tp1[h] += __shfl_xor_sync(0xffffffffu, tp1[h], 2);
          tp[h] += tp1[h];
        }
        if (qc == 0) {
          sT[qr] = fr[0] - tp[0];
          sT[8 + qr] = fr[1] - tp[1];
        }
        // ---- K(c+1,c+1) into the other scratch block (W_{c-1} has left it)
        if (c + 1 < ncol) {
          double* scn = sScr0 + ((c + 1) & 1) * BE;
          const double2 z = make_double2(0.0, 0.0);
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(scn)[lane + 32 * i] = z;
          __syncwarp();
#pragma unroll
          for (int i = 0; i < PREF; ++i)
            if (po1 + lane + 32 * i < pd1) scn[posd[i] & 255] = kvd[i];
          for (int q = po1 + 32 * PREF + lane; q < pd1; q += 32) scn[ldg_i32(a.b16_pos + q) & 255] = ldg_f64(kvs + q);
          if (lane < BT && (c + 1) * BT + lane >= a.n) scn[b16_off(lane, lane)] = 1.0;   // identity on the padded diagonal
        }
        __syncwarp();
        bar_sync_n(2, 64);                     // [G] -S is in the staging slots: P = -S + K
#pragma unroll
        for (int i = 0; i < PRET; ++i)
          if (po0 + lane + 32 * i < po1) {
            double* p = sm + sOff[poso[i] >> 8] + (poso[i] & 255);
            *p += kvo[i];
          }
        for (int q = po0 + 32 * PRET + lane; q < po1; q += 32) {
          const int pos = ldg_i32(a.b16_pos + q);
          double* p = sm + sOff[pos >> 8] + (pos & 255);
          *p += ldg_f64(kvs + q);
        }
      }
      __syncthreads();                         // [X] W_c is in the scratch block, P(c+rb, c) are in their slots
      BPH(4)
      fail = sOff[NB + 1];
      if (fail) break;   // uniform

      if (isG) {
        // ---------------- y_c = W_c t_c
        double yp[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double t4 = sT[ks * 4 + qc];
          yp[0] = fma(scr[fo(ks, lsw)], t4, yp[0]);
          yp[1] = fma(scr[fo(4 + ks, lsw)], t4, yp[1]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 1);
          yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 2);
        }
        if (qc == 0) {
          sY[yslot * BT + qr] = yp[0];
          sY[yslot * BT + 8 + qr] = yp[1];
          ysys[c * BT + qr] = yp[0];
          ysys[c * BT + 8 + qr] = yp[1];
        }
        po0 = pd1;
        po1 = pn1;
        pd1 = pn2 >= 0 ? pn2 : pn1;
      } else {
        // ---------------- L(c+rb, c) = P W^T into the ring: warp F takes rb = 1, warp T the rest (two separate code
        // paths, so that the chain warp's one block solve does not inherit the trailing warp's register pressure)
        double wf[2][4];
#pragma unroll
        for (int nbp = 0; nbp < 2; ++nbp)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) wf[nbp][ks] = scr[fo(nbp * 4 + ks, lsw)];
        auto solve_block = [&](double* blk) {
          double a4[2][4];
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) a4[mb][ks] = blk[fo(mb * 4 + ks, lsw)];
          __syncwarp();                        // P(rb) fully read before L(rb) overwrites it
          double x[2][2][2] = {};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int nbp = 0; nbp < 2; ++nbp)
                if (ks < 2 * nbp + 2) dmma(x[mb][nbp][0], x[mb][nbp][1], a4[mb][ks], wf[nbp][ks]);   // W is lower triangular
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nbp = 0; nbp < 2; ++nbp)
              *reinterpret_cast<double2*>(blk + cpair_off(mb, nbp, lane)) = make_double2(x[mb][nbp][0], x[mb][nbp][1]);
        };
        if (isF) {
          if ((nzc >> 1) & 1u) solve_block(sRing + idx[1] * BE);
        } else {
#pragma unroll
          for (int rb = 2; rb <= NB; ++rb)
            if ((nzc >> rb) & 1u) solve_block(sRing + (rb * (rb - 1) / 2 + idx[rb]) * BE);
        }
      }
      BPH(5)
      __syncthreads();                         // [Y] column c of L is in the ring, y_c in its slot
      BPH(6)

#pragma unroll
      for (int e = NB; e >= 2; --e) nzprev[e] = nzprev[e - 1];
      nzprev[1] = nzc;
#pragma unroll
      for (int e = 1; e <= NB; ++e) idx[e] = (idx[e] + 1 == e) ? 0 : idx[e] + 1;
      yslot = (yslot == NB) ? 0 : yslot + 1;
    }

    if (fail) {
      if (tid == 0) a.status[b] = fail;
      __syncthreads();
      continue;
    }
    if (isG) {                                 // the last block column of the factor -> HBM
      double* chunk = Lb + (int64_t)(ncol - 1) * (NB + 1) * BE;
      const double* wsrc = sScr0 + ((ncol - 1) & 1) * BE;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<double2*>(chunk)[lane + 32 * i] = reinterpret_cast<const double2*>(wsrc)[lane + 32 * i];
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nzprev[1] >> rb) & 1u)) continue;
        const int sl = idx[rb] == 0 ? rb - 1 : idx[rb] - 1;
        const double* src = sRing + (rb * (rb - 1) / 2 + sl) * BE;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<double2*>(chunk + rb * BE)[lane + 32 * i] = reinterpret_cast<const double2*>(src)[lane + 32 * i];
      }
    }
    __syncthreads();                           // the factor is in HBM/L2, the ring is free

    // ---------------- back substitution on warp F: u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb}), last block first
    if (isF) {
      auto fetch = [&](int c) {
        const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c) | 1u;     // bit 0: W_c
        const double* src = Lb + (int64_t)c * (NB + 1) * BE;
        double* dst = sRing + (c & 1) * (NB + 1) * BE;
#pragma unroll
        for (int e = 0; e <= NB; ++e)
          if ((nz >> e) & 1u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async16(dst + e * BE + (lane + 32 * i) * 2, src + e * BE + (lane + 32 * i) * 2);
          }
        cp_async_commit();
      };
      fetch(ncol - 1);
      // y_c and the block mask of column c are fetched one block column ahead (they come from L2)
      double ycn[4] = {0.0, 0.0, 0.0, 0.0};
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (ncol - 1) * BT + ks * 4 + qc);
      }
      unsigned nzb = (unsigned)ldg_i32(a.b16_nz + ncol - 1);
      for (int c = ncol - 1; c >= 0; --c) {
        yslot = (yslot == 0) ? NB : yslot - 1;
        double yc[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) yc[ks] = ycn[ks];
        const unsigned nz = nzb;
        if (c > 0) {
          nzb = (unsigned)ldg_i32(a.b16_nz + c - 1);
          if (qr == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ycn[ks] = __ldcg(ysys + (c - 1) * BT + ks * 4 + qc);
          }
        }
        if (c > 0) {
          fetch(c - 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const double* buf = sRing + (c & 1) * (NB + 1) * BE;
        double tp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int rb = 1; rb <= NB; ++rb) {
          if (!((nz >> rb) & 1u)) continue;
          int us = yslot + rb;
          if (us > NB) us -= NB + 1;
          const double* uv = sY + us * BT;
          const double* blk = buf + rb * BE;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double ur = uv[h * 8 + qr];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tp[ks] = fma(blk[fo(h * 4 + ks, lsw)], ur, tp[ks]);
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 4);
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 8);
          tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 16);
        }
        if (qr == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) sT[ks * 4 + qc] = yc[ks] - tp[ks];
        }
        __syncwarp();
        double up[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double rr = sT[h * 8 + qr];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) up[ks] = fma(buf[fo(h * 4 + ks, lsw)], rr, up[ks]);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 4);
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 8);
          up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 16);
        }
        if (qr == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            sY[yslot * BT + ks * 4 + qc] = up[ks];
            ysys[c * BT + ks * 4 + qc] = up[ks];
          }
        }
        __syncwarp();
      }
      BPH(7)
      if (lane == 0) a.status[b] = 0;
    }
    BPH_FLUSH(lane == 0 && isF)
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Load cases of one truss (LargeArgs::shared_k): the stiffness matrix is assembled and factorised once (system 0, by the
// kernels above) and every load case only runs the two substitutions against that factor, one warp per load case:
//   forward (right-looking)   y_c = W_c t_c,   t_{c+rb} -= L(c+rb,c) y_c        t = the load vector at the free DOFs
//   backward                  u_c = W_c^T (y_c - sum_rb L(c+rb,c)^T u_{c+rb})
// Both sweeps stream the column chunks [W_c | L(c+1,c) .. L(c+NB,c)] of the shared factor from L2 through a two-buffer
// cp.async pipeline.  Replaces the 1024 identical factorisations a loop over Truss.Solve() would do.
// ------------------------------------------------------------------------------------------------------------
template <int NB>
struct SubstCfg {
  static constexpr int DOUBLES = 2 * (NB + 1) * BE + (NB + 1) * BT + BT;   // two chunk buffers | t / u ring | block rhs
};

template <int NB, int NW>
__global__ void __launch_bounds__(32 * NW) k_band_subst(const LargeArgs a) {
  extern __shared__ __align__(16) double sm_all[];
  double* sm = sm_all + (threadIdx.x >> 5) * SubstCfg<NB>::DOUBLES;
  double* sBuf = sm;
  double* sY = sBuf + 2 * (NB + 1) * BE;       // ring: block c of t (then y), later of u, in slot c mod (NB+1)
  double* sT = sY + (NB + 1) * BT;
  const int lane = threadIdx.x & 31, lsw = lane_swz(lane);
  const int ncol = a.nb16;
  const int qr = lane >> 2, qc = lane & 3;
  const int st0 = a.status[0];                 // outcome of the shared factorisation
  const double* Lb = a.L;                      // factor of system 0

  for (int b = blockIdx.x * NW + (threadIdx.x >> 5); b < a.batch; b += gridDim.x * NW) {
    if (b > 0 && lane == 0) a.status[b] = st0;
    if (st0 != 0) continue;
    const double* fsys = a.force + b * a.force_stride;
    double* ysys = a.y + (int64_t)b * a.n_pad;
    auto fetch = [&](int c) {
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c) | 1u;     // bit 0: W_c
      const double* src = Lb + (int64_t)c * (NB + 1) * BE;
      double* dst = sBuf + (c & 1) * (NB + 1) * BE;
#pragma unroll
      for (int e = 0; e <= NB; ++e)
        if ((nz >> e) & 1u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async16(dst + e * BE + (lane + 32 * i) * 2, src + e * BE + (lane + 32 * i) * 2);
        }
      cp_async_commit();
    };
    auto load_rhs = [&](int c, int slot) {     // t_c = f at the free DOFs of block c (0 on the padding)
      if (lane < BT) {
        const int row = c * BT + lane;
        sY[slot * BT + lane] = (c < ncol && row < a.n) ? __ldg(fsys + __ldg(a.free_idx + row)) : 0.0;
      }
    };
    __syncwarp();
    fetch(0);
#pragma unroll
    for (int e = 0; e < NB; ++e) load_rhs(e, e);
    int slot = 0;                              // c mod (NB+1)
    // ---------------- forward sweep
    for (int c = 0; c < ncol; ++c) {
      {
        int sl = slot + NB;
        if (sl > NB) sl -= NB + 1;
        load_rhs(c + NB, sl);                  // block c+NB enters the window
      }
      if (c + 1 < ncol) {
        fetch(c + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c);
      const double* buf = sBuf + (c & 1) * (NB + 1) * BE;
      double yp[2] = {0.0, 0.0};               // y_c = W_c t_c
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const double t4 = sY[slot * BT + ks * 4 + qc];
        yp[0] = fma(buf[fo(ks, lsw)], t4, yp[0]);
        yp[1] = fma(buf[fo(4 + ks, lsw)], t4, yp[1]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 1);
        yp[h] += __shfl_xor_sync(0xffffffffu, yp[h], 2);
      }
      __syncwarp();
      if (qc == 0) {
        sT[qr] = yp[0];
        sT[8 + qr] = yp[1];
        ysys[c * BT + qr] = yp[0];
        ysys[c * BT + 8 + qr] = yp[1];
      }
      __syncwarp();
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {       // t_{c+rb} -= L(c+rb,c) y_c
        if (!((nz >> rb) & 1u)) continue;
        const double* blk = buf + rb * BE;
        double tq[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double y4 = sT[ks * 4 + qc];
          tq[0] = fma(blk[fo(ks, lsw)], y4, tq[0]);
          tq[1] = fma(blk[fo(4 + ks, lsw)], y4, tq[1]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tq[h] += __shfl_xor_sync(0xffffffffu, tq[h], 1);
          tq[h] += __shfl_xor_sync(0xffffffffu, tq[h], 2);
        }
        int sl = slot + rb;
        if (sl > NB) sl -= NB + 1;
        if (qc == 0) {
          sY[sl * BT + qr] -= tq[0];
          sY[sl * BT + 8 + qr] -= tq[1];
        }
      }
      __syncwarp();
      slot = (slot == NB) ? 0 : slot + 1;
    }
    // ---------------- backward sweep (the ring now collects u; block c in slot c mod (NB+1))
    fetch(ncol - 1);
    for (int c = ncol - 1; c >= 0; --c) {
      slot = (slot == 0) ? NB : slot - 1;
      double yc[4] = {0.0, 0.0, 0.0, 0.0};
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) yc[ks] = __ldcg(ysys + c * BT + ks * 4 + qc);
      }
      if (c > 0) {
        fetch(c - 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c);
      const double* buf = sBuf + (c & 1) * (NB + 1) * BE;
      double tp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nz >> rb) & 1u)) continue;
        int us = slot + rb;
        if (us > NB) us -= NB + 1;
        const double* uv = sY + us * BT;
        const double* blk = buf + rb * BE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double ur = uv[h * 8 + qr];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tp[ks] = fma(blk[fo(h * 4 + ks, lsw)], ur, tp[ks]);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 4);
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 8);
        tp[ks] += __shfl_xor_sync(0xffffffffu, tp[ks], 16);
      }
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) sT[ks * 4 + qc] = yc[ks] - tp[ks];
      }
      __syncwarp();
      double up[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double rr = sT[h * 8 + qr];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) up[ks] = fma(buf[fo(h * 4 + ks, lsw)], rr, up[ks]);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 4);
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 8);
        up[ks] += __shfl_xor_sync(0xffffffffu, up[ks], 16);
      }
      if (qr == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          sY[slot * BT + ks * 4 + qc] = up[ks];
          ysys[c * BT + ks * 4 + qc] = up[ks];
        }
      }
      __syncwarp();
    }
  }
}

// Eight load cases per warp on the FP64 tensor cores: the substitutions of a tile of load cases are 16x16 by 16x8
// block products (DMMA m8n8k4, the load cases in the n dimension), so one pass over the factor serves eight load cases
// (an eighth of k_band_subst's L2 traffic) and a block column costs a few dependent DMMA chains instead of shuffle
// reductions.  Columns of a DMMA product do not interact: a load case's result does not depend on its tile mates.
//   forward   Y_c = W_c T_c,  T_{c+rb} += (-L(c+rb,c)) Y_c      T (accumulator layout) lives in registers
//   backward  R = Y_c + sum_rb (-L(c+rb,c)^T) U_{c+rb},  U_c = W_c^T R
// Accumulator -> B-operand layout changes go through a 16x8 shared-memory scratch; all of y stays in shared memory.
template <int NB>
__global__ void __launch_bounds__(32) k_band_subst8(const LargeArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NBUF = 2;                      // chunk buffers [W_c | L(c+1,c) .. L(c+NB,c)]: one block column in flight (more did not help)
  double* sBuf = sm;
  double* sU = sBuf + NBUF * (NB + 1) * BE;       // ring of NB+1 blocks of u as [16][8] arrays (block c in slot c mod (NB+1))
  double* sX = sU + (NB + 1) * 128;            // [16][8] transpose scratch
  double* sYall = sX + 128;                    // [ncol][2][32][2] y, accumulator layout
  const int lane = threadIdx.x, lsw = lane_swz(lane);
  const int ncol = a.nb16;
  const int qr = lane >> 2, qc = lane & 3;
  const int st0 = a.status[0];                 // outcome of the shared factorisation
  const double* Lb = a.L;
  int tro[2][4];                               // element (4 ks + qc, 8 mb + qr) of a block: A operand of the transposed block
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tro[mb][ks] = b16_off(4 * ks + qc, 8 * mb + qr);

  for (int tile = blockIdx.x; tile * 8 < a.batch; tile += gridDim.x) {
    const int b0 = tile * 8;
    if (lane < 8 && b0 + lane < a.batch && b0 + lane > 0) a.status[b0 + lane] = st0;
    if (st0 != 0) continue;
    const int n0 = b0 + 2 * qc, n1 = n0 + 1;   // this lane's accumulator columns
    const bool v0 = n0 < a.batch, v1 = n1 < a.batch;
    const double* f0 = a.force + (int64_t)(v0 ? n0 : b0) * a.force_stride;
    const double* f1 = a.force + (int64_t)(v1 ? n1 : b0) * a.force_stride;
    auto fetch = [&](int c) {                  // out of range: an empty group, so the wait counts stay uniform
      if (c < 0 || c >= ncol) { cp_async_commit(); return; }
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c) | 1u;     // bit 0: W_c
      const double* src = Lb + (int64_t)c * (NB + 1) * BE;
      double* dst = sBuf + (c & (NBUF - 1)) * (NB + 1) * BE;
#pragma unroll
      for (int e = 0; e <= NB; ++e)
        if ((nz >> e) & 1u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async16(dst + e * BE + (lane + 32 * i) * 2, src + e * BE + (lane + 32 * i) * 2);
        }
      cp_async_commit();
    };
    auto load_idx = [&](int c, int (&fi)[2]) {      // free-DOF -> DOF index of rows 8 mb + qr of block c (-1 on the padding)
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int row = c * BT + 8 * mb + qr;
        fi[mb] = (c < ncol && row < a.n) ? __ldg(a.free_idx + row) : -1;
      }
    };
    auto load_t = [&](const int (&fi)[2], double (&t)[2][2]) {   // the load vectors at those rows
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        t[mb][0] = (v0 && fi[mb] >= 0) ? __ldg(f0 + fi[mb]) : 0.0;
        t[mb][1] = (v1 && fi[mb] >= 0) ? __ldg(f1 + fi[mb]) : 0.0;
      }
    };
    auto to_b = [&](const double (&x)[2][2], double (&bfrag)[4]) {   // accumulator layout -> B operand (k = row, n = load case)
      __syncwarp();
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
        *reinterpret_cast<double2*>(sX + (8 * mb + qr) * 8 + 2 * qc) = make_double2(x[mb][0], x[mb][1]);
      __syncwarp();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) bfrag[ks] = sX[(4 * ks + qc) * 8 + qr];
    };

    double t[NB + 1][2][2];
    int fin[2];                                // indices of the block that enters the window next: fetched a column ahead
#pragma unroll
    for (int e = 0; e <= NB; ++e) {
      load_idx(e, fin);
      load_t(fin, t[e]);
    }
    load_idx(NB + 1, fin);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i) fetch(i);
    // ---------------- forward sweep
    for (int c = 0; c < ncol; ++c) {
      double tn[2][2];
      load_t(fin, tn);                         // block c+NB+1 enters the window after this column; its loads fly meanwhile
      load_idx(c + NB + 2, fin);
      __syncwarp();                            // every lane is done with the buffer of column c-1, which takes column c+NBUF-1
      fetch(c + NBUF - 1);
      cp_async_wait<NBUF - 1>();
      __syncwarp();
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c);
      const double* buf = sBuf + (c & (NBUF - 1)) * (NB + 1) * BE;
      double bf[4];
      to_b(t[0], bf);
      double y[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) dmma(y[mb][0], y[mb][1], buf[fo(mb * 4 + ks, lsw)], bf[ks]);
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
        reinterpret_cast<double2*>(sYall)[(c * 2 + mb) * 32 + lane] = make_double2(y[mb][0], y[mb][1]);
      double yf[4];
      to_b(y, yf);
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nz >> rb) & 1u)) continue;
        const double* blk = buf + rb * BE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) dmma(t[rb][mb][0], t[rb][mb][1], -blk[fo(mb * 4 + ks, lsw)], yf[ks]);
      }
#pragma unroll
      for (int e = 0; e < NB; ++e)
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) { t[e][mb][0] = t[e + 1][mb][0]; t[e][mb][1] = t[e + 1][mb][1]; }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) { t[NB][mb][0] = tn[mb][0]; t[NB][mb][1] = tn[mb][1]; }
      __syncwarp();
    }
    // ---------------- backward sweep
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i) fetch(ncol - 1 - i);
    int slot = (ncol - 1) % (NB + 1);
    for (int c = ncol - 1; c >= 0; --c) {
      __syncwarp();
      fetch(c - (NBUF - 1));
      cp_async_wait<NBUF - 1>();
      __syncwarp();
      const unsigned nz = (unsigned)ldg_i32(a.b16_nz + c);
      const double* buf = sBuf + (c & (NBUF - 1)) * (NB + 1) * BE;
      double r[2][2][2];                       // two partial sums (odd / even rb) keep the dependent DMMA chains short
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const double2 v = reinterpret_cast<const double2*>(sYall)[(c * 2 + mb) * 32 + lane];
        r[0][mb][0] = v.x;
        r[0][mb][1] = v.y;
        r[1][mb][0] = r[1][mb][1] = 0.0;
      }
#pragma unroll
      for (int rb = 1; rb <= NB; ++rb) {
        if (!((nz >> rb) & 1u)) continue;
        int us = slot + rb;
        if (us > NB) us -= NB + 1;
        const double* ub = sU + us * 128;
        const double* blk = buf + rb * BE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double bu = ub[(4 * ks + qc) * 8 + qr];
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) dmma(r[rb & 1][mb][0], r[rb & 1][mb][1], -blk[tro[mb][ks]], bu);
        }
      }
      double rs[2][2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        rs[mb][0] = r[0][mb][0] + r[1][mb][0];
        rs[mb][1] = r[0][mb][1] + r[1][mb][1];
      }
      double rf[4];
      to_b(rs, rf);
      double u[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) dmma(u[mb][0], u[mb][1], buf[tro[mb][ks]], rf[ks]);
      double* ub = sU + slot * 128;
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        *reinterpret_cast<double2*>(ub + (8 * mb + qr) * 8 + 2 * qc) = make_double2(u[mb][0], u[mb][1]);
        const int row = c * BT + 8 * mb + qr;
        if (v0) a.y[(int64_t)n0 * a.n_pad + row] = u[mb][0];
        if (v1) a.y[(int64_t)n1 * a.n_pad + row] = u[mb][1];
      }
      __syncwarp();
      slot = (slot == 0) ? NB : slot - 1;
    }
  }
}

template <int NB>
int launch_subst(const LargeArgs& a, int num_sm, cudaStream_t st) {
  static const int force = [] { const char* s = getenv("TB_SUBST_TILE"); return s ? atoi(s) : 0; }();   // 1 | 8: force a kernel
  const int smem8 = (2 * (NB + 1) * BE + (NB + 1) * 128 + 128 + a.nb16 * 128) * 8;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  dev &= 63;                                   // (attribute caches are per device: cudaFuncSetAttribute is a per-device setting)
  tb_prof_begin(TB_PROF_SUBST, st);
  if (force != 1 && smem8 <= 200 * 1024) {
    // tiles of eight load cases on the tensor cores; y of the whole system stays in shared memory
    static int granted_dev[64] = {};
    int& granted = granted_dev[dev];
    if (granted < smem8) {
      cudaError_t e = cudaFuncSetAttribute(k_band_subst8<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8);
      if (e != cudaSuccess) return (int)e;
      granted = smem8;
    }
    const int tiles = (a.batch + 7) / 8, cap = num_sm * (int)((220 * 1024) / (smem8 + 1024));
    k_band_subst8<NB><<<tiles < cap ? tiles : cap, 32, smem8, st>>>(a);
  } else {
    constexpr int NW = 4;
    const int smem = NW * SubstCfg<NB>::DOUBLES * 8;
    static bool set_dev[64] = {};
    bool& set = set_dev[dev];
    if (!set) {
      cudaError_t e = cudaFuncSetAttribute(k_band_subst<NB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      set = true;
    }
    k_band_subst<NB, NW><<<(a.batch + NW - 1) / NW, 32 * NW, smem, st>>>(a);
  }
  tb_prof_end(TB_PROF_SUBST, st);
  return (int)cudaGetLastError();
}

template <int NB>
int launch_band(const LargeArgs& a, int num_sm, cudaStream_t st) {
  // one warp per system when the batch alone fills the GPU (fewest instructions per system), else several warps per
  // system: TB_BAND_WARPS = 1 | 2 | 3 forces a kernel (parity tests, comparisons)
  static const int force = [] { const char* s = getenv("TB_BAND_WARPS"); return s ? atoi(s) : 0; }();
  // k_band1 packs independent systems (one warp each) into a CTA: four while four rings fit the 227 KB of an SM's
  // shared memory (NB <= 6), two for the widest bands (NB = 7, 8: 59 / 76 KB per system)
  constexpr int NW = NB <= 6 ? 4 : 2;
  const int smem1 = NW * BandCfg<NB>::DOUBLES * 8, smem2 = (BandCfg<NB>::DOUBLES + 208) * 8, smem3 = Band3Cfg<NB>::DOUBLES * 8;
  static int per_dev[64][3] = {};              // attribute / occupancy queries once per instantiation and device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  int &per1 = per_dev[dev & 63][0], &per2 = per_dev[dev & 63][1], &per3 = per_dev[dev & 63][2];
  if (per1 == 0) {
    int q1 = 0, q2 = 0, q3 = 0;
    cudaError_t e = cudaFuncSetAttribute(k_band1<NB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_band2<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_band3<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q1, k_band1<NB, NW>, 32 * NW, smem1);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q2, k_band2<NB>, 64, smem2);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q3, k_band3<NB>, 96, smem3);
    if (e != cudaSuccess) return (int)e;
    per3 = q3 < 1 ? 1 : q3;
    per2 = q2 < 1 ? 1 : q2;
    per1 = q1 < 1 ? 1 : q1;
    if (getenv("TB_BAND_DEBUG")) fprintf(stderr, "launch_band<%d>: CTAs/SM k_band1 %d, k_band2 %d, k_band3 %d\n", NB, q1, q2, q3);
  }
  // NB > 5: the trailing warp's accumulators spill -> one warp per system.  Otherwise three warps per system (96
  // registers per thread, six systems per SM) while the batch fits in one wave of that kernel, else two warps per
  // system (eight systems per SM; measured faster than k_band1's eight one-warp systems per SM at every batch size:
  // bar-942 x8192 in 2.73 ms against 2.96 ms)
  int warps = NB > 5 ? 1 : (int64_t)a.batch <= (int64_t)num_sm * per3 ? 3 : 2;
  if (NB <= 5 && a.band_warps == 2) warps = 2;
  if (force == 1 || (NB <= 5 && (force == 2 || force == 3))) warps = force;
  tb_prof_begin(TB_PROF_CHOL, st);
  if (warps == 3) {
    int grid = num_sm * per3;
    if (grid > a.batch) grid = a.batch;
    k_band3<NB><<<grid, 96, smem3, st>>>(a);
  } else if (warps == 2) {
    int grid = num_sm * per2;
    if (grid > a.batch) grid = a.batch;
    k_band2<NB><<<grid, 64, smem2, st>>>(a);
  } else {
    int grid = num_sm * per1;
    if (grid > (a.batch + NW - 1) / NW) grid = (a.batch + NW - 1) / NW;
    k_band1<NB, NW><<<grid, 32 * NW, smem1, st>>>(a);
  }
  tb_prof_end(TB_PROF_CHOL, st);
  return (int)cudaGetLastError();
}

}  // namespace

int tb_band_smem_bytes(int NB) {
  const int ring = NB * (NB + 1) / 2, buf = ring > 2 * (NB + 1) ? ring : 2 * (NB + 1);
  return (buf * BE + BE + (NB + 1) * BT + BT + 32 + 8) * 8;
}

int tb_launch_band_subst(const LargeArgs& a, int num_sm, cudaStream_t st) {
  switch (a.NB) {
    case 1: return launch_subst<1>(a, num_sm, st);
    case 2: return launch_subst<2>(a, num_sm, st);
    case 3: return launch_subst<3>(a, num_sm, st);
    case 4: return launch_subst<4>(a, num_sm, st);
    case 5: return launch_subst<5>(a, num_sm, st);
    case 6: return launch_subst<6>(a, num_sm, st);
    case 7: return launch_subst<7>(a, num_sm, st);
    case 8: return launch_subst<8>(a, num_sm, st);
    default: return TB_ERR_TOO_LARGE;
  }
}

int tb_launch_band_chol(const LargeArgs& a, int num_sm, cudaStream_t st) {
  switch (a.NB) {
    case 1: return launch_band<1>(a, num_sm, st);
    case 2: return launch_band<2>(a, num_sm, st);
    case 3: return launch_band<3>(a, num_sm, st);
    case 4: return launch_band<4>(a, num_sm, st);
    case 5: return launch_band<5>(a, num_sm, st);
    case 6: return launch_band<6>(a, num_sm, st);
    case 7: return launch_band<7>(a, num_sm, st);
    case 8: return launch_band<8>(a, num_sm, st);
    default: return TB_ERR_TOO_LARGE;
  }
}
