// Two-sided fused band kernel (program: tb_ts.cuh / tb_tsplan.cu).  One CTA of two warps solves one system:
//
//   warp 0 ("top")     eliminates the block columns of T top-down, then the separator S, in its own band numbering
//   warp 1 ("bottom")  eliminates the block columns of B bottom-up (a plain Cholesky in its reversed numbering) and hands
//                      its Schur-complement contribution to S over to warp 0
//
// so the dependent pivot chain of a system is max(|T|, |B|) + |S| instead of n (bar-942: 376 instead of 696).
// Per block column c of a side (8 columns, the FP64 MMA shape, so no padding inside blocks):
//
//   products   S(c+rb, c) = sum_d L(c+rb, c-d) L(c, c-d)^T        DMMA m8n8k4; operands from a ring of the live 8x8 blocks
//                                                                 in shared memory (operand-fragment layout), accumulators
//                                                                 in registers; the forward substitution rides on the
//                                                                 B-operand registers
//   assemble   K(c+rb, c)  from the member products k (c_i c_j)   Member.matK / Truss.GetKMatrix, truss.py:65-86,307-316:
//                                                                 the members of this block column are recomputed from the
//                                                                 joint positions (one lane per member, the reference's
//                                                                 roundings), every K entry sums its contributions in
//                                                                 ascending member order -- K never exists in HBM
//   factor     rows [P(c,c); I; P(c+1,c); t_c^T] (lanes = rows)   column elimination in LDL^T form: the chain per pivot is
//                                                                 1/d (seed + one cubic step) and one FMA; column values
//                                                                 and the next pivot's ingredients travel by shuffles
//                                                                 one step ahead, the scaling by rsqrt(d) is off the chain.
//                                                                 Yields L_D, Z = L_D^{-T}, L(c+1,c) and y_c at once
//   solves     L(c+rb, c) = P(c+rb, c) Z for rb >= 2              DMMA
//   stores     [Z | y_c | L(c+rb, c)] as one contiguous chunk     read back by the back substitution with one
//                                                                 cp.async.bulk (TMA) per block column, three chunks in flight
//
// The back substitution runs per side from the separator outwards; u goes to the workspace in the internal order and the
// recovery kernel (tb_large.cu, k_recover) turns it into displacements, member forces and reactions.
// Replaces np.linalg.solve (slientruss3d/truss.py:343) for narrow-band systems.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "tb_common.cuh"
#include "tb_blocks.cuh"
#include "tb_ts.cuh"

#ifdef TB_PHASE_TIMING
__device__ unsigned long long g_ts_cycles[16];
#define TPH_DECL long long _ph_t = clock64(); unsigned long long _ph_acc[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0};
#define TPH(i) { long long _n = clock64(); _ph_acc[i] += (unsigned long long)(_n - _ph_t); _ph_t = _n; }
#define TPH_FLUSH(cond) if (cond) { for (int _i = 0; _i < 16; ++_i) atomicAdd(&g_ts_cycles[_i], _ph_acc[_i]); }
extern "C" int tb_ts_phase_read(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, g_ts_cycles, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_ts_cycles, z, sizeof(z));
  return 0;
}
#else
#define TPH_DECL
#define TPH(i) {}
#define TPH_FLUSH(cond)
extern "C" int tb_ts_phase_read(unsigned long long* out) {
  for (int i = 0; i < 16; ++i) out[i] = 0;
  return 0;
}
#endif

namespace {

// (not volatile: the scheduler may sink an MMA below later shared-memory loads, so operand loads run ahead of the tensor pipe)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
using tbblk::rsqrt_pos;

constexpr int NSTAGE = TS_NSTAGE;
constexpr int X_SCR = TS_X_SCR, X_T = TS_X_T, X_Y = TS_X_Y, X_MISC = TS_X_MISC, X_TOTAL = TS_X_TOTAL;
__device__ __forceinline__ int b8_off(int r, int k) { return ts_b8_off(r, k); }

// 1/d for a normal positive d: hardware seed (relative error ~2^-20) and one cubic step, three dependent FP64 operations
__device__ __forceinline__ double rcp_pos(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TS_DONE;\n"
      "bra TS_WAIT;\n"
      "TS_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one contiguous chunk global -> shared through the TMA unit; completion is signalled on the mbarrier (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// Column elimination of the rows [P(c,c) (lanes 0-7) | I (8-15) | P(c+1,c) (16-23) | t^T (24)]: every lane owns one row.
// With d_k the k-th pivot and s_k = 1/d_k, step k subtracts (row[k] s_k) a_jk from row[j], j > k, where a_jk is entry k of
// row j of P (lane j).  What a step needs from other lanes -- the column a_jk, j > k, and the ingredients p, q of the next
// pivot d_{k+1} = p - q s_k -- are values of the state BEFORE the step's own update is known, so they are shuffled while
// the reciprocal is computed: the dependent chain per pivot is one reciprocal and one FMA.  All rows end up scaled by
// rsqrt(d_k) per column (computed beside the chain): lanes 0-7 hold L_D, 8-15 Z = L_D^{-T}, 16-23 L(c+1,c) = P(c+1,c) Z,
// lane 24 y^T = t^T Z.  Returns 0 or k+1 for the first non-positive pivot.
__device__ __forceinline__ int factor_rows8(double (&row)[8], int lane) {
  const unsigned FULL = 0xffffffffu;
  int bad = 0;
  double rs[8];
  double d = __shfl_sync(FULL, row[0], 0);
  double colv[8];
#pragma unroll
  for (int j = 1; j < 8; ++j) colv[j] = __shfl_sync(FULL, row[0], j);
  double p = __shfl_sync(FULL, row[1], 1);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!(d > 1e-290) && !bad) bad = k + 1;
    const double s = rcp_pos(d);
    rs[k] = rsqrt_pos(d);
    double dn = 0.0;
    if (k < 7) dn = fma(-(colv[k + 1] * colv[k + 1]), s, p);     // next pivot: needs only s
    const double t = row[k] * s;
#pragma unroll
    for (int j = k + 1; j < 8; ++j) row[j] = fma(-t, colv[j], row[j]);
    if (k < 7) {
#pragma unroll
      for (int j = k + 2; j < 8; ++j) colv[j] = __shfl_sync(FULL, row[k + 1], j);
      if (k < 6) p = __shfl_sync(FULL, row[k + 2], k + 2);
      d = dn;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) row[k] *= rs[k];
  return bad;
}

// shared-memory accesses by 32-bit address (the ring pointers are kept as byte addresses with the lane offset folded in)
__device__ __forceinline__ double lds64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// Operand loads of the products: not volatile, so that the scheduler can issue them well ahead of the MMAs.  Safe there:
// nothing writes the ring between the previous column's last __syncwarp() and the stores of this column's staging, and
// those stores take the accumulators these loads feed (a data dependence keeps every load in front of them).
__device__ __forceinline__ double lds64_nv(unsigned addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64_32_nv(unsigned addr) {
  double v;
  asm("ld.shared.f64 %0, [%1+32];" : "=d"(v) : "r"(addr));
  return v;
}

// The block products of a column as one flat list: pair i -> (distance d, block row rb), d = 1..NB, rb = 0..NB-d
template <int NB>
__host__ __device__ constexpr int pair_d(int i) {
  int d = 1;
  while (i >= NB - d + 1) { i -= NB - d + 1; ++d; }
  return d;
}
template <int NB>
__host__ __device__ constexpr int pair_rb(int i) {
  int d = 1;
  while (i >= NB - d + 1) { i -= NB - d + 1; ++d; }
  return i;
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// NB: sub-diagonal blocks of the wider side's band view (loops, the pointer ring and the accumulators are sized by it)
template <int NB>
__global__ void __launch_bounds__(64, 7) k_band_ts(const TsArgs a) {
  extern __shared__ __align__(16) double sm_all[];
  TPH_DECL
  const unsigned FULL = 0xffffffffu;
  // (the warp index through a shuffle: the compiler then knows that everything derived from it is warp-uniform)
  const int side = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const TsSideDev& S = a.side[side];
  const int ncol_own = S.ncol_own, ncol_tot = S.ncol_tot, nbs = S.nb;
  const bool two = a.side[1].ncol_tot > 0;
  const int main0 = ts_main_doubles(a.side[0].nb, a.chunk_max), main1 = two ? ts_main_doubles(a.side[1].nb, a.chunk_max) : 0;
  double* sm = sm_all + (side ? main0 + X_TOTAL : 0);
  const int mainsz = side ? main1 : main0;
  double* sRing = sm;
  double* sScr = sm + mainsz + X_SCR;
  double* sT = sm + mainsz + X_T;
  double* sY = sm + mainsz + X_Y;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sm + mainsz + X_MISC);           // NSTAGE mbarriers
  int* sFlag = reinterpret_cast<int*>(sm_all + main0 + X_MISC + 10);            // CTA-wide: [0] pivot failure
  double* sUS = sm_all + main0 + X_SCR;                                         // separator displacements (top -> bottom), top's scratch block

  const int qr = lane >> 2, qc = lane & 3;
  const int cpo = ((qc >> 1) << 5) + ((qr ^ (qc & 2)) << 2) + ((qc & 1) << 1);  // this lane's accumulator pair inside a block (ts_b8_off)
  const unsigned sl1 = (lane & 8) ? 192u : 320u;                                // from this lane's operand element in slab 0 to the one in slab 1 (rows r ^ 2)
  const int nS = a.nS;
  const unsigned sm_u32 = smem_u32(sm);
  const unsigned ring_u32 = sm_u32 + lane * 8;                                  // byte address of this lane's operand element in slot 0
  const unsigned scr_u32 = smem_u32(sScr);

  if (lane == 0) {                                                              // (side 1's area exists even when it has no columns)
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&sBar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;                                                           // bit i: parity the next wait on stage i expects
  __syncthreads();

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (threadIdx.x == 0) sFlag[0] = 0;
    __syncthreads();
    const double* fsys = a.force + (int64_t)b * a.force_stride;
    double* Lsys = a.L + (int64_t)b * a.l_per_sys;
    double* Xsys = a.X + (int64_t)b * nS * nS * TS_BE;
    double* Zsys = a.Z + (int64_t)b * nS * TS_BT;
    double* ufs = a.uf + (int64_t)b * a.n_pad;
    const int instat = a.status[b];                                             // input problem flagged by the assembly pass

    // Ring of the live blocks: Q[e][j], 1 <= j <= e, is the address of block (c-j+e, c-j) (diagonal e, created j columns
    // ago; it dies after column c-j+e).  At column c the B operand of distance d is Q[d][d], the A operand of block row rb
    // is Q[rb+d][d], and the new block (c+e, c) takes over the slot of the dying block (c, c-e) = Q[e][e]: the pointers
    // rotate by register moves, no index arithmetic anywhere (and the slot sequence is ts_ring_slot, known to the plan).
    unsigned Q[NB + 1][NB + 1];
#pragma unroll
    for (int e = 1; e <= NB; ++e)
#pragma unroll
      for (int j = 1; j <= e; ++j) Q[e][j] = ring_u32 + (unsigned)((e * (e - 1) / 2 + (j - 1)) * TS_BE * 8);
    unsigned Yq[NB + 1];                                                        // Yq[d]: address of y_{c-d}[qc] (d >= 1); Yq[0]: the slot y_c will take
#pragma unroll
    for (int e = 0; e <= NB; ++e) Yq[e] = smem_u32(sY) + (unsigned)(e * TS_BT + qc) * 8u;
    int fail = 0;

    // running pointers into the program: one record per block column, the rhs rows, the K values and their positions
    const int4* recp = S.colrec;
    const int32_t* dofp = S.rowdof + qr;
    const double* kvp = a.kv + (int64_t)b * a.nnz + S.ent0 + lane;
    const int32_t* epp = a.epos + S.ent0 + lane;
    int4 recn = make_int4(0, 0, 0, 0);
    int dof_n = -1;
    if (ncol_tot > 0) {
      recn = __ldg(recp);
      dof_n = __ldg(dofp);
    }
    TPH(0)

    // =========================================== factorisation + forward substitution
    for (int c = 0; c < ncol_tot; ++c) {
      const bool own = c < ncol_own;
      const bool xcol = side == 1 && !own;                                      // bottom side, separator column: products only
      if (side == 0 && c == ncol_own && two) pair_sync(1);                      // the bottom side's hand-over is complete
      const unsigned nzc = (unsigned)recn.x & 511u, xm = ((unsigned)recn.x >> 9) & 511u;
      const int ecnt = (int)((unsigned)recn.x >> 18), lof = recn.y;
      const unsigned long long pmask = (unsigned)recn.z | ((unsigned long long)(unsigned)recn.w << 32);   // live block products
      const int dof_c = dof_n;
      const double fr = (!xcol && dof_c >= 0) ? __ldg(fsys + dof_c) : 0.0;
      // K values of this block column (assembly pass, program order): in flight while the products run
      double kvr[TS_EPL];
      int kpos[TS_EPL];
#pragma unroll
      for (int i = 0; i < TS_EPL; ++i) {
        kvr[i] = 0.0;
        kpos[i] = -1;
        if (lane + 32 * i < ecnt) {
          kvr[i] = __ldg(kvp + 32 * i);
          kpos[i] = __ldg(epp + 32 * i);
        }
      }
      const double* kvx = kvp + 32 * TS_EPL;                                    // (entries beyond the prefetched ones: rare)
      const int32_t* epx = epp + 32 * TS_EPL;
      kvp += ecnt;
      epp += ecnt;
      if (c + 1 < ncol_tot) {
        recp += 1;
        recn = __ldg(recp);
        dofp += TS_BT;
        dof_n = __ldg(dofp);
      }
      TPH(1)

      // ---------------- products with the previous nb block columns
      double acc[NB + 2][2];
#pragma unroll
      for (int rb = 0; rb <= NB + 1; ++rb) acc[rb][0] = acc[rb][1] = 0.0;
      double tp = 0.0;
      // (a flat, fully predicated software pipeline over all NB(NB+1)/2 products was tried: 8 % faster for a lone system,
      // 3-5 % slower with seven systems per SM -- the dead products' predicated-off instructions still take issue slots)
      {
        int base = 0;                                                           // flat index of (d, rb = 0)
#pragma unroll
        for (int d = 1; d <= NB; ++d) {
          const unsigned bits = (unsigned)(pmask >> base);                      // bit rb: product (d, rb) is live
          base += NB - d + 1;
          if (!(bits & 1u)) continue;                                           // L(c, c-d) structurally zero (uniform)
          double b0 = lds64_nv(Q[d][d]), b1 = lds64_nv(Q[d][d] + sl1);
          {
            const double y0 = lds64_nv(Yq[d]), y1 = lds64_32_nv(Yq[d]);
            tp = fma(b0, y0, tp);
            tp = fma(b1, y1, tp);
          }
          b0 = -b0;                                                            // the accumulators collect -S: P = K - S needs no pass of its own
          b1 = -b1;
          // block rows two at a time: the second k-slab of one block issues behind the first k-slab of the other
#pragma unroll
          for (int rb = 0; rb + d <= NB; rb += 2) {
            const int e = rb + d;
            const bool on0 = (bits >> rb) & 1u, on1 = e + 1 <= NB && ((bits >> (rb + 1)) & 1u);
            const int e1i = e + 1 <= NB ? e + 1 : e;
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            if (on0) { a00 = lds64_nv(Q[e][d]); a01 = lds64_nv(Q[e][d] + sl1); }
            if (on1) { a10 = lds64_nv(Q[e1i][d]); a11 = lds64_nv(Q[e1i][d] + sl1); }
            if (on0) dmma(acc[rb][0], acc[rb][1], a00, b0);
            if (on1) dmma(acc[rb + 1][0], acc[rb + 1][1], a10, b0);
            if (on0) dmma(acc[rb][0], acc[rb][1], a01, b1);
            if (on1) dmma(acc[rb + 1][0], acc[rb + 1][1], a11, b1);
          }
        }
      }
      tp += __shfl_xor_sync(FULL, tp, 1);
      tp += __shfl_xor_sync(FULL, tp, 2);
      __syncwarp();
      TPH(2)

      if (xcol) {
        // ---------------- hand-over of the bottom side: block (c+rb, c) here is separator block (row J+rb, column J) of the
        // top side, J = nS-1-jq-rb, transposed and flipped: element (r, k) -> (7-k, 7-r)
        const int jq = c - ncol_own;
#pragma unroll
        for (int rb = 0; rb <= NB; ++rb) {
          if (!((nzc >> rb) & 1u)) continue;
          double* dst = Xsys + (int64_t)((nS - 1 - jq - rb) * nS + rb) * TS_BE;
          dst[(7 - 2 * qc) * 8 + (7 - qr)] = -acc[rb][0];
          dst[(6 - 2 * qc) * 8 + (7 - qr)] = -acc[rb][1];
        }
        if (qc == 0) Zsys[(nS - 1 - jq) * TS_BT + (7 - qr)] = tp;
      } else {
        if (side == 0 && !own && two) {
          const int J = c - ncol_own;
#pragma unroll
          for (int rb = 0; rb <= NB; ++rb) {
            if (!((xm >> rb) & 1u)) continue;
            const double2 x = __ldcg(reinterpret_cast<const double2*>(Xsys + (int64_t)(J * nS + rb) * TS_BE + qr * 8 + 2 * qc));
            acc[rb][0] -= x.x;
            acc[rb][1] -= x.y;
          }
          tp += __ldcg(Zsys + J * TS_BT + qr);
        }
        // ---------------- P = K - S: -S into the staging slots (block rb: slot of the dying block (c, c-rb); rb = 0: scratch),
        // then every K value is added at its position (resolved by the plan: byte offset from this side's base)
#pragma unroll
        for (int rb = 0; rb <= NB; ++rb) {
          if (!((nzc >> rb) & 1u)) continue;
          const unsigned ad = (rb == 0 ? scr_u32 : Q[rb][rb] - lane * 8) + cpo * 8;
          asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ad), "d"(acc[rb][0]), "d"(acc[rb][1]) : "memory");
        }
        if (qc == 0) sT[qr] = fr - tp;
        __syncwarp();
        {   // (the positions of a block column are distinct: all loads, then all stores)
          double pv[TS_EPL];
#pragma unroll
          for (int i = 0; i < TS_EPL; ++i) {
            pv[i] = 0.0;
            if (kpos[i] >= 0) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(pv[i]) : "r"(sm_u32 + (unsigned)kpos[i]));
          }
#pragma unroll
          for (int i = 0; i < TS_EPL; ++i)
            if (kpos[i] >= 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sm_u32 + (unsigned)kpos[i]), "d"(pv[i] + kvr[i]) : "memory");
        }
        for (int e = 32 * TS_EPL + lane; e < ecnt; e += 32) {                   // rare: more than 32 * TS_EPL entries in this block column
          const int kp = __ldg(epx + (e - 32 * TS_EPL - lane));
          if (kp < 0) continue;
          const unsigned ad = sm_u32 + (unsigned)kp;
          double v;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ad));
          v += __ldg(kvx + (e - 32 * TS_EPL - lane));
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad), "d"(v) : "memory");
        }
        if (qc == 0 && dof_c < 0) sScr[b8_off(qr, qr)] = 1.0;                   // identity on padding (no entries there)
        __syncwarp();
        TPH(3)

        // ---------------- rows: P(c,c) | I | P(c+1,c) | t^T
        const bool has1 = (nzc >> 1) & 1u;
        const unsigned slot1 = Q[1][1] - lane * 8;                              // block (c+1, c)
        double row[8];
        {
          const unsigned blk = lane < 8 ? smem_u32(sScr) : slot1;
          const bool ld = lane < 8 || (lane >= 16 && lane < 24 && has1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const unsigned src = blk + h * 256 + (((lane & 7) ^ (2 * h)) << 5);   // row (lane & 7), columns 4h .. 4h+3
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            if (ld) {
              asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(src));
              asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v2), "=d"(v3) : "r"(src + 16));
            }
            row[4 * h] = v0; row[4 * h + 1] = v1; row[4 * h + 2] = v2; row[4 * h + 3] = v3;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (lane >= 8 && lane < 16) row[k] = (k == lane - 8) ? 1.0 : 0.0;
            if (lane == 24) row[k] = sT[k];
          }
        }
        __syncwarp();
        TPH(4)
        {
          const int badk = factor_rows8(row, lane);
          if (badk && !fail) {
            const int nat = __ldg(S.rownat + c * TS_BT + badk - 1);
            fail = (nat >= 0 ? nat : 0) + 1;
          }
        }
        TPH(5)

        // ---------------- Z (as the solves' operand, and to HBM), L(c+1,c) into the ring and to HBM, y_c
        double* chunk = Lsys + lof;
        if (lane >= 8 && lane < 16) {
          const int i = lane - 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) sScr[b8_off(k, i)] = row[k];             // W[k][i] = Z[i][k]
#pragma unroll
          for (int k = 0; k < 8; ++k) chunk[k * 8 + i] = row[k];                 // the chunk holds Z^T (the back substitution's lanes read it without bank conflicts)
        } else if (lane >= 16 && lane < 24) {
          if (has1) {
            const int i = lane - 16;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(slot1 + h * 256 + ((i ^ (2 * h)) << 5)), "d"(row[4 * h]), "d"(row[4 * h + 1]) : "memory");
              asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(slot1 + h * 256 + ((i ^ (2 * h)) << 5) + 16), "d"(row[4 * h + 2]), "d"(row[4 * h + 3]) : "memory");
            }
#pragma unroll
            for (int h = 0; h < 4; ++h)
              reinterpret_cast<double2*>(chunk + TS_BE + TS_BT + i * 8)[h] = make_double2(row[2 * h], row[2 * h + 1]);
          }
        } else if (lane == 24) {                                                 // (qc == 0: Yq[0] is the base of the new y slot)
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(Yq[0] + h * 16), "d"(row[2 * h]), "d"(row[2 * h + 1]) : "memory");
            reinterpret_cast<double2*>(chunk + TS_BE)[h] = make_double2(row[2 * h], row[2 * h + 1]);
          }
        }
        __syncwarp();
        const double w0 = sScr[lane], w1 = sScr[32 + (lane ^ 8)];

        // ---------------- solves L(c+rb, c) = P(c+rb, c) Z, rb >= 2
        {
          double f0[NB + 1], f1[NB + 1], x0[NB + 1], x1[NB + 1];
#pragma unroll
          for (int rb = 2; rb <= NB; ++rb) {
            f0[rb] = f1[rb] = 0.0;
            if (!((nzc >> rb) & 1u)) continue;
            f0[rb] = lds64(Q[rb][rb]);
            f1[rb] = lds64(Q[rb][rb] + sl1);
          }
          __syncwarp();                                                          // every P block is read before any L overwrites it
#pragma unroll
          for (int rb = 2; rb <= NB; ++rb) {
            x0[rb] = x1[rb] = 0.0;
            if (!((nzc >> rb) & 1u)) continue;
            dmma(x0[rb], x1[rb], f0[rb], w0);
          }
          int rank = (nzc >> 1) & 1u;
#pragma unroll
          for (int rb = 2; rb <= NB; ++rb) {
            if (!((nzc >> rb) & 1u)) continue;
            dmma(x0[rb], x1[rb], f1[rb], w1);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(Q[rb][rb] - lane * 8 + cpo * 8), "d"(x0[rb]), "d"(x1[rb]) : "memory");
            *reinterpret_cast<double2*>(chunk + TS_BE + TS_BT + rank * TS_BE + qr * 8 + 2 * qc) = make_double2(x0[rb], x1[rb]);
            ++rank;
          }
        }
        __syncwarp();
        TPH(6)
      }

      // ---------------- advance the ring: the slot of the dying block of every diagonal becomes the youngest block's
#pragma unroll
      for (int e = 1; e <= NB; ++e) {
        const unsigned dying = Q[e][e];
#pragma unroll
        for (int j = e; j >= 2; --j) Q[e][j] = Q[e][j - 1];
        Q[e][1] = dying;
      }
      {   // ... and the y slots: the oldest becomes the next column's
        const unsigned oldest = Yq[NB];
#pragma unroll
        for (int j = NB; j >= 1; --j) Yq[j] = Yq[j - 1];
        Yq[0] = oldest;
      }
    }
    if (fail && lane == 0) atomicCAS(&sFlag[0], 0, fail);
    if (side == 1 && two) {
      __threadfence_block();
      pair_sync(1);                                                              // hand-over written (pairs with the top side's wait)
    }
    TPH(7)

    // =========================================== back substitution
    // u_c = Z_c (y_c - sum_rb L(c+rb, c)^T u_{c+rb}); the chunks come back through cp.async.bulk, NSTAGE in flight
    const int ncb = side == 0 ? ncol_tot : ncol_own;
    double* sBuf = sRing;
    double* sU = sRing + NSTAGE * a.chunk_max;                                   // ring of the last nb+1 blocks of u: block c in slot c mod (nb+1)
    fence_proxy_async();                                                         // this lane's factor stores / ring stores before the async proxy
    __syncwarp();
    auto issue = [&](int c, int stage) {
      const int4 rc = __ldg(S.colrec + c);
      const unsigned bytes = (unsigned)(TS_BE + TS_BT + TS_BE * __popc(((unsigned)rc.x >> 1) & 255u)) * 8u;
      mbar_expect_tx(&sBar[stage], bytes);
      bulk_g2s(sBuf + stage * a.chunk_max, Lsys + rc.y, bytes, &sBar[stage]);
    };
    if (lane == 0)
      for (int i = 0; i < NSTAGE; ++i)
        if (ncb - 1 - i >= 0) issue(ncb - 1 - i, i);
    const int g = lane >> 3, col = lane & 7;
    double uprev = 0.0;
    // the last nb+1 blocks of u live in a ring addressed by rotating pointers: Uq[d] = address of u_{c+d}[2g] (d >= 1),
    // Uq[0] = the slot u_c takes (all with this lane's 2g offset folded in)
    const unsigned su_u32 = smem_u32(sU);
    unsigned Uq[NB + 1];
#pragma unroll
    for (int e = 0; e <= NB; ++e) Uq[e] = su_u32 + (unsigned)(e * TS_BT + 2 * g) * 8u;
    if (side == 1 && two) {
      pair_sync(2);                                                              // separator displacements are published
      // virtual separator column ncol_own + j is u_{c+1+j} of the first back-substituted column c = ncol_own - 1: slot Uq[1+j]
      const int cnt = nS < nbs ? nS : nbs;                                       // (the band reaches nb separator blocks at most)
      if (lane < TS_BT) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (j < cnt) {
            const double v = sUS[(nS - 1 - j) * TS_BT + (7 - lane)];
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(Uq[1 + j] - (unsigned)(2 * g) * 8u + (unsigned)lane * 8u), "d"(v) : "memory");
          }
      }
      __syncwarp();
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(uprev) : "r"(Uq[1] - (unsigned)(2 * g) * 8u + (unsigned)col * 8u));
    }
    unsigned maskn = ncb > 0 ? ((unsigned)__ldg(&S.colrec[ncb - 1].x) & 511u) : 0u;
    int natn = ncb > 0 ? __ldg(S.rownat + (ncb - 1) * TS_BT + col) : -1;
    int stage = 0;
    const unsigned buf0 = smem_u32(sBuf);
    const unsigned lane_l = (unsigned)((2 * g) * 8 + col) * 8u, lane_z = lane_l;      // L[2g][col], L[2g+1][col]; Z^T[2g][col], Z^T[2g+1][col]
    for (int c = ncb - 1; c >= 0; --c) {
      const unsigned mask = maskn;
      const int nat = natn;
      if (c > 0) {
        maskn = ((unsigned)__ldg(&S.colrec[c - 1].x) & 511u);
        natn = __ldg(S.rownat + (c - 1) * TS_BT + col);
      }
      mbar_wait(&sBar[stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
      const unsigned bufa = buf0 + (unsigned)(stage * a.chunk_max) * 8u;
      // t[col] = sum_rb sum_r L(c+rb,c)[r][col] u_{c+rb}[r]; this lane: r = 2g, 2g+1.  The blocks further away first (their u is
      // in the ring), the neighbour last (its u has just been produced and comes by shuffle)
      double t = 0.0, t2 = 0.0;
      unsigned la = bufa + (unsigned)(TS_BE + TS_BT) * 8u + lane_l + (((mask >> 1) & 1u) ? (unsigned)TS_BE * 8u : 0u);
#pragma unroll
      for (int rb = 2; rb <= NB; ++rb) {
        if (!((mask >> rb) & 1u)) continue;
        double l0, l1, u0, u1;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l0) : "r"(la));
        asm volatile("ld.shared.f64 %0, [%1+64];" : "=d"(l1) : "r"(la));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(u0) : "r"(Uq[rb]));
        asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(u1) : "r"(Uq[rb]));
        la += (unsigned)TS_BE * 8u;
        if (rb & 1) { t = fma(l0, u0, t); t = fma(l1, u1, t); }
        else { t2 = fma(l0, u0, t2); t2 = fma(l1, u1, t2); }
      }
      double yc, z0, z1, l10 = 0.0, l11 = 0.0;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(yc) : "r"(bufa + (unsigned)(TS_BE + col) * 8u));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(z0) : "r"(bufa + lane_z));
      asm volatile("ld.shared.f64 %0, [%1+64];" : "=d"(z1) : "r"(bufa + lane_z));
      if ((mask >> 1) & 1u) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l10) : "r"(bufa + (unsigned)(TS_BE + TS_BT) * 8u + lane_l));
        asm volatile("ld.shared.f64 %0, [%1+64];" : "=d"(l11) : "r"(bufa + (unsigned)(TS_BE + TS_BT) * 8u + lane_l));
      }
      t += t2;
      {
        const double u0 = __shfl_sync(FULL, uprev, 2 * g), u1 = __shfl_sync(FULL, uprev, 2 * g + 1);
        t = fma(l10, u0, t);
        t = fma(l11, u1, t);
      }
      t += __shfl_xor_sync(FULL, t, 8);
      t += __shfl_xor_sync(FULL, t, 16);
      const double rr = yc - t;
      const double r0 = __shfl_sync(FULL, rr, 2 * g), r1 = __shfl_sync(FULL, rr, 2 * g + 1);
      double u = z0 * r0;
      u = fma(z1, r1, u);
      u += __shfl_xor_sync(FULL, u, 8);
      u += __shfl_xor_sync(FULL, u, 16);
      uprev = u;
      if (g == 0) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(Uq[0] + (unsigned)col * 8u), "d"(u) : "memory");   // (g == 0: no 2g offset in Uq)
        if (nat >= 0) ufs[nat] = u;
        if (side == 0 && c >= ncol_own) sUS[(c - ncol_own) * TS_BT + col] = u;
      }
      __syncwarp();
      if (lane == 0 && c - NSTAGE >= 0) issue(c - NSTAGE, stage);
      if (side == 0 && two && c == ncol_own) pair_sync(2);
      stage = stage + 1 == NSTAGE ? 0 : stage + 1;
      {   // u_c becomes u_{(c-1)+1}: rotate the ring pointers, the oldest slot is the next column's
        const unsigned oldest = Uq[NB];
#pragma unroll
        for (int j = NB; j >= 1; --j) Uq[j] = Uq[j - 1];
        Uq[0] = oldest;
      }
    }
    TPH(8)
    __syncthreads();
    if (threadIdx.x == 0 && instat == 0) a.status[b] = sFlag[0];
    __syncthreads();
  }
  TPH_FLUSH(lane == 0)
}

struct DevCache {
  int smem_set[TS_NBX + 1] = {};
};
DevCache g_cache[64];

}  // namespace

int tb_ts_smem_bytes(const TsPlan* ts) {
  const int m0 = ts_main_doubles(ts->side[0].nb, ts->chunk_max);
  const int m1 = ts->side[1].ncol_tot > 0 ? ts_main_doubles(ts->side[1].nb, ts->chunk_max) : 0;
  return (m0 + X_TOTAL + m1 + X_TOTAL) * 8;
}

template <int NB>
int launch_nb(const TsArgs& a, int smem, int grid, cudaStream_t st) {
  int dev = 0;
  TB_CUDA(cudaGetDevice(&dev));
  auto kern = k_band_ts<NB>;
  DevCache& dc = g_cache[dev & 63];
  if (dc.smem_set[NB] < smem) {
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // seven CTAs of ~30 KB per SM need the largest shared-memory carve-out
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dc.smem_set[NB] = smem;
  }
  kern<<<grid, 64, smem, st>>>(a);
  return 0;
}

int tb_launch_band_ts(const TsArgs& a, int smem, int num_sm, cudaStream_t st) {
  if (a.batch <= 0) return 0;
  if (num_sm <= 0) num_sm = 148;
  int per_sm = (228 * 1024) / (smem + 1024);
  if (per_sm > 7) per_sm = 7;
  if (per_sm < 1) per_sm = 1;
  const int grid = a.batch < num_sm * per_sm ? a.batch : num_sm * per_sm;
  const int nbm = a.side[0].nb > a.side[1].nb ? a.side[0].nb : a.side[1].nb;
  tb_prof_begin(TB_PROF_CHOL, st);
  int rc = 0;
  switch (nbm) {
    case 1: rc = launch_nb<1>(a, smem, grid, st); break;
    case 2: rc = launch_nb<2>(a, smem, grid, st); break;
    case 3: rc = launch_nb<3>(a, smem, grid, st); break;
    case 4: rc = launch_nb<4>(a, smem, grid, st); break;
    case 5: rc = launch_nb<5>(a, smem, grid, st); break;
    case 6: rc = launch_nb<6>(a, smem, grid, st); break;
    case 7: rc = launch_nb<7>(a, smem, grid, st); break;
    default: rc = launch_nb<8>(a, smem, grid, st); break;
  }
  tb_prof_end(TB_PROF_CHOL, st);
  if (rc) return rc;
  tb_count_launch(1);
  return (int)cudaGetLastError();
}
