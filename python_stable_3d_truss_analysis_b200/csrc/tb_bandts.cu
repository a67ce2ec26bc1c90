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

using tbblk::dmma;
using tbblk::rsqrt_pos;

constexpr int NBX = TS_NBX;
constexpr int NSTAGE = 3;                        // factor chunks in flight during the back substitution
// per side, after the ring: [64 diagonal staging, then Z as a DMMA operand | 4*TS_CHUNK member (k, c) | 8 rhs | 16 misc]
constexpr int X_SCR = 0, X_PROD = TS_BE, X_T = X_PROD + 4 * TS_CHUNK, X_MISC = X_T + TS_BT, X_TOTAL = X_MISC + 16;

__host__ __device__ inline int ts_main_doubles(int nb, int chunk_max) {
  const int ring = nb * (nb + 1) / 2 * TS_BE;
  const int back = NSTAGE * chunk_max + (nb + 1) * TS_BT;
  return ring > back ? ring : back;
}

// element (r, k) of an 8x8 block in operand-fragment layout: slab k / 4 holds [row 8][k 4]
__device__ __forceinline__ int b8_off(int r, int k) { return ((k >> 2) << 5) + (r << 2) + (k & 3); }

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

template <int DIM>
__global__ void __maxnreg__(144) k_band_ts(const TsArgs a) {
  extern __shared__ __align__(16) double sm_all[];
  TPH_DECL
  const unsigned FULL = 0xffffffffu;
  const int side = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TsSideDev& S = a.side[side];
  const int nb = S.nb;
  const bool two = a.side[1].ncol_tot > 0;
  const int main0 = ts_main_doubles(a.side[0].nb, a.chunk_max), main1 = two ? ts_main_doubles(a.side[1].nb, a.chunk_max) : 0;
  double* sm = sm_all + (side ? main0 + X_TOTAL : 0);
  const int mainsz = side ? main1 : main0;
  double* sRing = sm;
  double* sScr = sm + mainsz + X_SCR;
  double* sProd = sm + mainsz + X_PROD;
  double* sT = sm + mainsz + X_T;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sm + mainsz + X_MISC);           // NSTAGE mbarriers
  int* sOff = reinterpret_cast<int*>(sm + mainsz + X_MISC + 4);                 // [NBX+1] staging slot of block rb (doubles from sm)
  int* sFlag = reinterpret_cast<int*>(sm_all + main0 + X_MISC + 10);            // CTA-wide: [0] pivot failure, [1] input problem
  double* sUS = sm_all + main0 + X_PROD;                                        // separator displacements (top -> bottom), top's product area

  const int qr = lane >> 2, qc = lane & 3;
  const int cpo = ((qc >> 1) << 5) + (qr << 2) + ((qc & 1) << 1);               // this lane's accumulator pair inside a block
  const int nS = a.nS;

  if (lane == 0) {                                                              // (side 1's area exists even when it has no columns)
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&sBar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;                                                           // bit i: parity the next wait on stage i expects
  __syncthreads();

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (threadIdx.x == 0) { sFlag[0] = 0; sFlag[1] = 0; }
    __syncthreads();
    const double* xyz = a.xyz + (int64_t)b * a.xyz_stride;
    const double* fsys = a.force + (int64_t)b * a.force_stride;
    double* Lsys = a.L + (int64_t)b * a.l_per_sys;
    double* Xsys = a.X + (int64_t)b * nS * nS * TS_BE;
    double* Zsys = a.Z + (int64_t)b * nS * TS_BT;
    double* ufs = a.uf + (int64_t)b * a.n_pad;
    double* kdbg = a.kdebug ? a.kdebug + (int64_t)b * a.kdbg_stride + (side ? a.kdbg_off1 : 0) : nullptr;

    int idx[NBX + 1];
    unsigned nzprev[NBX + 1];
    double yreg[NBX][2];
#pragma unroll
    for (int e = 0; e <= NBX; ++e) { idx[e] = 0; nzprev[e] = 0u; }
#pragma unroll
    for (int e = 0; e < NBX; ++e) yreg[e][0] = yreg[e][1] = 0.0;
    int gflag = 0, fail = 0;

    // ---- raw inputs of the member this lane handles in the first chunk of a block column, fetched one column ahead
    double rx0[DIM], rx1[DIM], rar = 0.0, re = 0.0;
    bool rok = true;
    const int4 none4 = make_int4(-1, 0, 0, 0);
    auto load_raw = [&](const int4& dsc) {
      rok = true;
      rar = re = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i) rx0[i] = rx1[i] = 0.0;
      if (dsc.x < 0) return;
      if (a.gene) {
        const int g = __ldg(a.gene + (int64_t)b * a.gene_stride + dsc.x);
        if ((unsigned)g < (unsigned)a.n_type) {
          rar = __ldg(a.type_table + 3 * g);
          re = __ldg(a.type_table + 3 * g + 1);
        } else {
          rok = false;
        }
      } else {
        const double* t = a.aed + (int64_t)b * a.aed_stride + 3 * (int64_t)dsc.x;
        rar = __ldg(t);
        re = __ldg(t + 1);
      }
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        rx0[i] = __ldg(xyz + dsc.y * DIM + i);
        rx1[i] = __ldg(xyz + dsc.z * DIM + i);
      }
    };
    // member products of the reference (truss.py:19,56-63): k = e a / L, c_i = dx_i / L, with its roundings
    auto geometry = [&](int slot, bool live) {
      if (!live) return;
      double dx[DIM], cc[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) dx[i] = __dsub_rn(rx1[i], rx0[i]);
      double l2 = __dmul_rn(dx[0], dx[0]);
#pragma unroll
      for (int i = 1; i < DIM; ++i) l2 = __dadd_rn(l2, __dmul_rn(dx[i], dx[i]));
      const double len = __dsqrt_rn(l2);
      double k = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i) cc[i] = 0.0;
      if (!rok) {
        gflag = min(gflag, TB_INFO_BAD_INDEX);
      } else if (!(len > 0.0)) {
        gflag = min(gflag, TB_INFO_ZERO_LENGTH);
      } else {
        k = __ddiv_rn(__dmul_rn(re, rar), len);
#pragma unroll
        for (int i = 0; i < DIM; ++i) cc[i] = __ddiv_rn(dx[i], len);
      }
      double* o = sProd + slot * 4;
      o[0] = k;
#pragma unroll
      for (int i = 0; i < DIM; ++i) o[1 + i] = cc[i];
    };
    // one contribution: +-k (c_i c_j), product first, then the scale (truss.py:69-70 / 80-81)
    auto term = [&](int pk) {
      const double* o = sProd + (pk >> 4) * 4;
      const int ij = pk & 7;
      int i, j;
      if (DIM == 3) {
        i = (ij >= 3) + (ij >= 5);
        j = ij < 3 ? ij : (ij < 5 ? ij - 2 : 2);
      } else {
        i = ij >= 2;
        j = ij >= 1;
      }
      const double t = __dmul_rn(o[0], __dmul_rn(o[1 + i], o[1 + j]));
      return (pk & 8) ? -t : t;
    };

    // software pipeline over the block columns: masks / entry ranges / rhs row one column ahead, member descriptor two
    // columns ahead, the member's raw inputs one column ahead (every address depends on c only)
    int4 cin = none4, cen = make_int4(0, 0, 0, 0), mdn = none4;
    int dof_n = -1;
    bool live_n = false;
    if (S.ncol_tot > 0) {
      cin = __ldg(S.colinfo);
      cen = __ldg(S.colent);
      dof_n = __ldg(S.rowdof + qr);
      mdn = __ldg(S.mem0 + lane);
      live_n = mdn.x >= 0;
      load_raw(mdn);
      mdn = S.ncol_tot > 1 ? __ldg(S.mem0 + TS_CHUNK + lane) : none4;
    }
    TPH(0)

    // =========================================== factorisation + forward substitution
    for (int c = 0; c < S.ncol_tot; ++c) {
      const bool own = c < S.ncol_own;
      const bool xcol = side == 1 && !own;                                      // bottom side, separator column: products only
      if (side == 0 && c == S.ncol_own && two) pair_sync(1);                    // the bottom side's hand-over is complete
      const unsigned nzc = (unsigned)cin.x, srcc = (unsigned)cin.y, xm = (unsigned)cin.z;
      const double fr = (!xcol && dof_n >= 0) ? __ldg(fsys + dof_n) : 0.0;
      // member products of this block column (first chunk) from the inputs fetched during the previous column
      int e0 = cen.x, e1 = cen.y;
      const int q0 = cen.z, q1 = cen.w;
      if (q1 > q0) geometry(lane, live_n);
      int2 ed[TS_EPL];
#pragma unroll
      for (int i = 0; i < TS_EPL; ++i) {
        ed[i] = make_int2(0, 0);
        if (e0 + lane + 32 * i < e1) ed[i] = __ldg(S.ent + e0 + lane + 32 * i);
      }
      if (c + 1 < S.ncol_tot) {
        cin = __ldg(S.colinfo + c + 1);
        cen = __ldg(S.colent + c + 1);
        dof_n = __ldg(S.rowdof + (c + 1) * TS_BT + qr);
      }
      live_n = mdn.x >= 0;
      load_raw(mdn);                                                            // (no member: nothing is loaded)
      mdn = c + 2 < S.ncol_tot ? __ldg(S.mem0 + (c + 2) * TS_CHUNK + lane) : none4;
      TPH(1)

      // ---------------- products with the previous nb block columns
      double acc[NBX + 1][2];
#pragma unroll
      for (int rb = 0; rb <= NBX; ++rb) acc[rb][0] = acc[rb][1] = 0.0;
      double tp = 0.0;
#pragma unroll
      for (int d = 1; d <= NBX; ++d) {
        if (d > nb) break;
        const unsigned nzp = nzprev[d];
        if (!((nzp >> d) & 1u)) continue;                                       // L(c, c-d) structurally zero (uniform)
        const double* Bm = sRing + (d * (d - 1) / 2 + idx[d]) * TS_BE;
        const double b0 = Bm[lane], b1 = Bm[32 + lane];
        tp = fma(b0, yreg[d - 1][0], tp);
        tp = fma(b1, yreg[d - 1][1], tp);
        double a0[NBX + 1], a1[NBX + 1];
#pragma unroll
        for (int rb = 0; rb + d <= NBX; ++rb) {
          const int e = rb + d;
          a0[rb] = a1[rb] = 0.0;
          if (e > nb || !((nzp >> e) & 1u)) continue;
          int sl = idx[e] - d;
          if (sl < 0) sl += e;
          const double* A = sRing + (e * (e - 1) / 2 + sl) * TS_BE;
          a0[rb] = A[lane];
          a1[rb] = A[32 + lane];
        }
#pragma unroll
        for (int rb = 0; rb + d <= NBX; ++rb) {
          const int e = rb + d;
          if (e > nb || !((nzp >> e) & 1u)) continue;
          dmma(acc[rb][0], acc[rb][1], a0[rb], b0);
        }
#pragma unroll
        for (int rb = 0; rb + d <= NBX; ++rb) {
          const int e = rb + d;
          if (e > nb || !((nzp >> e) & 1u)) continue;
          dmma(acc[rb][0], acc[rb][1], a1[rb], b1);
        }
      }
      tp += __shfl_xor_sync(FULL, tp, 1);
      tp += __shfl_xor_sync(FULL, tp, 2);
      __syncwarp();
      TPH(2)

      if (xcol) {
        // ---------------- hand-over of the bottom side: block (c+rb, c) here is separator block (row J+rb, column J) of the
        // top side, J = nS-1-jq-rb, transposed and flipped: element (r, k) -> (7-k, 7-r)
        const int jq = c - S.ncol_own;
#pragma unroll
        for (int rb = 0; rb <= NBX; ++rb) {
          if (rb > nb || !((nzc >> rb) & 1u)) continue;
          double* dst = Xsys + (int64_t)((nS - 1 - jq - rb) * nS + rb) * TS_BE;
          dst[(7 - 2 * qc) * 8 + (7 - qr)] = acc[rb][0];
          dst[(6 - 2 * qc) * 8 + (7 - qr)] = acc[rb][1];
        }
        if (qc == 0) Zsys[(nS - 1 - jq) * TS_BT + (7 - qr)] = tp;
#pragma unroll
        for (int e = NBX - 1; e >= 1; --e) { yreg[e][0] = yreg[e - 1][0]; yreg[e][1] = yreg[e - 1][1]; }
        yreg[0][0] = yreg[0][1] = 0.0;
      } else {
        if (side == 0 && !own && two) {
          const int J = c - S.ncol_own;
#pragma unroll
          for (int rb = 0; rb <= NBX; ++rb) {
            if (rb > nb || !((xm >> rb) & 1u)) continue;
            const double2 x = __ldcg(reinterpret_cast<const double2*>(Xsys + (int64_t)(J * nS + rb) * TS_BE + qr * 8 + 2 * qc));
            acc[rb][0] += x.x;
            acc[rb][1] += x.y;
          }
          tp += __ldcg(Zsys + J * TS_BT + qr);
        }
        // ---------------- stage K(:,c): block rb into the slot of the dead block (c, c-rb); rb = 0 into the scratch block
        if (lane <= nb) sOff[lane] = lane == 0 ? (int)(sScr - sm) : (lane * (lane - 1) / 2 + idx[lane]) * TS_BE;
        {
          const double2 z = make_double2(0.0, 0.0);
          reinterpret_cast<double2*>(sScr)[lane] = z;
#pragma unroll
          for (int rb = 1; rb <= NBX; ++rb)
            if (rb <= nb && ((nzc >> rb) & 1u)) reinterpret_cast<double2*>(sRing + (rb * (rb - 1) / 2 + idx[rb]) * TS_BE)[lane] = z;
        }
        __syncwarp();
        for (int q = q0; q < q1; ++q) {
          if (q > q0) {                        // further chunks of a crowded block column: fetched on the spot
            const int m0 = __ldg(S.mem_ptr + q), m1 = __ldg(S.mem_ptr + q + 1);
            int4 dsc = make_int4(-1, 0, 0, 0);
            if (m0 + lane < m1) dsc = __ldg(S.mem + m0 + lane);
            // the prefetched inputs of the next column are live in the r* registers: save and restore them
            const double sar = rar, se = re;
            const bool sok = rok;
            double sx0[DIM], sx1[DIM];
#pragma unroll
            for (int i = 0; i < DIM; ++i) { sx0[i] = rx0[i]; sx1[i] = rx1[i]; }
            load_raw(dsc);
            __syncwarp();
            geometry(lane, dsc.x >= 0);
            rar = sar; re = se; rok = sok;
#pragma unroll
            for (int i = 0; i < DIM; ++i) { rx0[i] = sx0[i]; rx1[i] = sx1[i]; }
            e0 = __ldg(S.ent_ptr + q);
            e1 = __ldg(S.ent_ptr + q + 1);
          }
          __syncwarp();
          for (int e = e0 + lane, i = 0; e < e1; e += 32, ++i) {
            int2 dsc;
            if (q == q0 && i < TS_EPL) {
              dsc = ed[0];
#pragma unroll
              for (int u = 1; u < TS_EPL; ++u) dsc = i == u ? ed[u] : dsc;
            } else {
              dsc = __ldg(S.ent + e);
            }
            const int pos = dsc.x & 1023, cnt = dsc.x >> 10;
            double* dst = sm + sOff[pos >> 6] + (pos & 63);
            double v = *dst;
            if (cnt == 1) {
              v = __dadd_rn(v, term(dsc.y));
            } else {
              for (int t = 0; t < cnt; t += 4) {           // four map loads in flight, summed in ascending member order
                int pk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pk[u] = t + u < cnt ? __ldg(S.pack + dsc.y + t + u) : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (pk[u] >= 0) v = __dadd_rn(v, term(pk[u]));
              }
            }
            *dst = v;
            if (kdbg) kdbg[e] = v;
          }
          __syncwarp();
        }
        if (lane < TS_BT && __ldg(S.rowdof + c * TS_BT + lane) < 0) sScr[b8_off(lane, lane)] = 1.0;   // identity on padding
        __syncwarp();
        TPH(3)

        // ---------------- P = K - S in place (accumulator pairs), right-hand side of the block
#pragma unroll
        for (int rb = 0; rb <= NBX; ++rb) {
          if (rb > nb || !((nzc >> rb) & 1u)) continue;
          double* blk = rb == 0 ? sScr : sRing + (rb * (rb - 1) / 2 + idx[rb]) * TS_BE;
          double2* pp = reinterpret_cast<double2*>(blk + cpo);
          double2 v = *pp;
          v.x -= acc[rb][0];
          v.y -= acc[rb][1];
          *pp = v;
        }
        if (qc == 0) sT[qr] = fr - tp;
        __syncwarp();

        // ---------------- rows: P(c,c) | I | P(c+1,c) | t^T
        double row[8];
        {
          const bool has1 = nb >= 1 && ((nzc >> 1) & 1u);
          const double* src = lane < 8 ? sScr : sRing + idx[1] * TS_BE;        // ring slot of diagonal 1: block (c+1, c)
          const int r = lane & 7;
          const bool ld = lane < 8 || (lane >= 16 && lane < 24 && has1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double2 v0 = make_double2(0.0, 0.0), v1 = v0;
            if (ld) {
              v0 = *reinterpret_cast<const double2*>(src + h * 32 + r * 4);
              v1 = *reinterpret_cast<const double2*>(src + h * 32 + r * 4 + 2);
            }
            row[4 * h] = v0.x; row[4 * h + 1] = v0.y; row[4 * h + 2] = v1.x; row[4 * h + 3] = v1.y;
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
        double* chunk = Lsys + __ldg(S.lofs + c);
        if (lane >= 8 && lane < 16) {
          const int i = lane - 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) sScr[b8_off(k, i)] = row[k];             // W[k][i] = Z[i][k]
#pragma unroll
          for (int h = 0; h < 4; ++h) reinterpret_cast<double2*>(chunk + i * 8)[h] = make_double2(row[2 * h], row[2 * h + 1]);
        } else if (lane >= 16 && lane < 24) {
          if (nb >= 1 && ((nzc >> 1) & 1u)) {
            const int i = lane - 16;
            double* blk = sRing + idx[1] * TS_BE;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              *reinterpret_cast<double2*>(blk + h * 32 + i * 4) = make_double2(row[4 * h], row[4 * h + 1]);
              *reinterpret_cast<double2*>(blk + h * 32 + i * 4 + 2) = make_double2(row[4 * h + 2], row[4 * h + 3]);
            }
#pragma unroll
            for (int h = 0; h < 4; ++h)
              reinterpret_cast<double2*>(chunk + TS_BE + TS_BT + i * 8)[h] = make_double2(row[2 * h], row[2 * h + 1]);
          }
        } else if (lane == 24) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            reinterpret_cast<double2*>(sT)[h] = make_double2(row[2 * h], row[2 * h + 1]);
            reinterpret_cast<double2*>(chunk + TS_BE)[h] = make_double2(row[2 * h], row[2 * h + 1]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int e = NBX - 1; e >= 1; --e) { yreg[e][0] = yreg[e - 1][0]; yreg[e][1] = yreg[e - 1][1]; }
        yreg[0][0] = sT[qc];
        yreg[0][1] = sT[4 + qc];
        const double w0 = sScr[lane], w1 = sScr[32 + lane];

        // ---------------- solves L(c+rb, c) = P(c+rb, c) Z, rb >= 2
        {
          double f0[NBX + 1], f1[NBX + 1], x0[NBX + 1], x1[NBX + 1];
#pragma unroll
          for (int rb = 2; rb <= NBX; ++rb) {
            f0[rb] = f1[rb] = 0.0;
            if (rb > nb || !((nzc >> rb) & 1u)) continue;
            const double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * TS_BE;
            f0[rb] = blk[lane];
            f1[rb] = blk[32 + lane];
          }
          __syncwarp();                                                          // every P block is read before any L overwrites it
#pragma unroll
          for (int rb = 2; rb <= NBX; ++rb) {
            x0[rb] = x1[rb] = 0.0;
            if (rb > nb || !((nzc >> rb) & 1u)) continue;
            dmma(x0[rb], x1[rb], f0[rb], w0);
          }
          int rank = (nzc >> 1) & 1u;
#pragma unroll
          for (int rb = 2; rb <= NBX; ++rb) {
            if (rb > nb || !((nzc >> rb) & 1u)) continue;
            dmma(x0[rb], x1[rb], f1[rb], w1);
            double* blk = sRing + (rb * (rb - 1) / 2 + idx[rb]) * TS_BE;
            *reinterpret_cast<double2*>(blk + cpo) = make_double2(x0[rb], x1[rb]);
            *reinterpret_cast<double2*>(chunk + TS_BE + TS_BT + rank * TS_BE + qr * 8 + 2 * qc) = make_double2(x0[rb], x1[rb]);
            ++rank;
          }
        }
        __syncwarp();
        TPH(6)
      }

      // ---------------- advance the ring
#pragma unroll
      for (int e = NBX; e >= 2; --e) nzprev[e] = nzprev[e - 1];
      nzprev[1] = srcc;
#pragma unroll
      for (int e = 1; e <= NBX; ++e) idx[e] = (idx[e] + 1 == e) ? 0 : idx[e] + 1;
    }
    if (fail && lane == 0) atomicCAS(&sFlag[0], 0, fail);
    if (gflag) atomicMin(&sFlag[1], gflag);
    if (side == 1 && two) {
      __threadfence_block();
      pair_sync(1);                                                              // hand-over written (pairs with the top side's wait)
    }
    TPH(7)

    // =========================================== back substitution
    // u_c = Z_c (y_c - sum_rb L(c+rb, c)^T u_{c+rb}); the chunks come back through cp.async.bulk, NSTAGE in flight
    const int ncb = side == 0 ? S.ncol_tot : S.ncol_own;
    const int ur = nb + 1;
    double* sBuf = sRing;
    double* sU = sRing + NSTAGE * a.chunk_max;                                   // ring of the last nb+1 blocks of u: block c in slot c mod (nb+1)
    fence_proxy_async();                                                         // this lane's factor stores / ring stores before the async proxy
    __syncwarp();
    auto issue = [&](int c, int stage) {
      const int o0 = __ldg(S.lofs + c), o1 = __ldg(S.lofs + c + 1);
      const unsigned bytes = (unsigned)(o1 - o0) * 8u;
      mbar_expect_tx(&sBar[stage], bytes);
      bulk_g2s(sBuf + stage * a.chunk_max, Lsys + o0, bytes, &sBar[stage]);
    };
    if (lane == 0)
      for (int i = 0; i < NSTAGE; ++i)
        if (ncb - 1 - i >= 0) issue(ncb - 1 - i, i);
    const int g = lane >> 3, col = lane & 7;
    double uprev = 0.0;
    int cs = ncb > 0 ? (ncb - 1) % ur : 0;                                       // ring slot of block c
    if (side == 1 && two) {
      pair_sync(2);                                                              // separator displacements are published
      int s2 = S.ncol_own % ur;
      const int cend = S.ncol_own + (nS < nb ? nS : nb);                          // (the band reaches nb separator blocks at most)
      for (int c = S.ncol_own; c < cend; ++c) {
        if (lane < TS_BT) sU[s2 * TS_BT + lane] = sUS[(nS - 1 - (c - S.ncol_own)) * TS_BT + (7 - lane)];
        s2 = s2 + 1 == ur ? 0 : s2 + 1;
      }
      __syncwarp();
      uprev = sU[(S.ncol_own % ur) * TS_BT + col];
    }
    unsigned maskn = ncb > 0 ? (unsigned)__ldg(&S.colinfo[ncb - 1].x) : 0u;
    int natn = ncb > 0 ? __ldg(S.rownat + (ncb - 1) * TS_BT + col) : -1;
    int stage = 0;
    for (int c = ncb - 1; c >= 0; --c) {
      const unsigned mask = maskn;
      const int nat = natn;
      if (c > 0) {
        maskn = (unsigned)__ldg(&S.colinfo[c - 1].x);
        natn = __ldg(S.rownat + (c - 1) * TS_BT + col);
      }
      mbar_wait(&sBar[stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
      const double* buf = sBuf + stage * a.chunk_max;
      // t[col] = sum_rb sum_r L(c+rb,c)[r][col] u_{c+rb}[r]; this lane: r = 2g, 2g+1.  The blocks further away first (their u is
      // in the ring), the neighbour last (its u has just been produced and comes by shuffle)
      double t = 0.0;
      int rank = (mask >> 1) & 1u;
#pragma unroll
      for (int rb = 2; rb <= NBX; ++rb) {
        if (rb > nb || !((mask >> rb) & 1u)) continue;
        const double* blk = buf + TS_BE + TS_BT + rank * TS_BE;
        ++rank;
        int us = cs + rb;
        if (us >= ur) us -= ur;
        t = fma(blk[(2 * g) * 8 + col], sU[us * TS_BT + 2 * g], t);
        t = fma(blk[(2 * g + 1) * 8 + col], sU[us * TS_BT + 2 * g + 1], t);
      }
      {
        const double u0 = __shfl_sync(FULL, uprev, 2 * g), u1 = __shfl_sync(FULL, uprev, 2 * g + 1);
        if (nb >= 1 && ((mask >> 1) & 1u)) {
          const double* blk = buf + TS_BE + TS_BT;
          t = fma(blk[(2 * g) * 8 + col], u0, t);
          t = fma(blk[(2 * g + 1) * 8 + col], u1, t);
        }
      }
      t += __shfl_xor_sync(FULL, t, 8);
      t += __shfl_xor_sync(FULL, t, 16);
      const double rr = buf[TS_BE + col] - t;
      const double r0 = __shfl_sync(FULL, rr, 2 * g), r1 = __shfl_sync(FULL, rr, 2 * g + 1);
      double u = buf[col * 8 + 2 * g] * r0;
      u = fma(buf[col * 8 + 2 * g + 1], r1, u);
      u += __shfl_xor_sync(FULL, u, 8);
      u += __shfl_xor_sync(FULL, u, 16);
      uprev = u;
      if (g == 0) {
        sU[cs * TS_BT + col] = u;
        if (nat >= 0) ufs[nat] = u;
        if (side == 0 && c >= S.ncol_own) sUS[(c - S.ncol_own) * TS_BT + col] = u;
      }
      __syncwarp();
      if (lane == 0 && c - NSTAGE >= 0) issue(c - NSTAGE, stage);
      if (side == 0 && two && c == S.ncol_own) pair_sync(2);
      stage = stage + 1 == NSTAGE ? 0 : stage + 1;
      cs = cs == 0 ? ur - 1 : cs - 1;
    }
    TPH(8)
    __syncthreads();
    if (threadIdx.x == 0) a.status[b] = sFlag[1] ? sFlag[1] : sFlag[0];
    __syncthreads();
  }
  TPH_FLUSH(lane == 0)
}

struct DevCache {
  int smem_set[2] = {0, 0};
};
DevCache g_cache[64];

}  // namespace

int tb_ts_smem_bytes(const TsPlan* ts) {
  const int m0 = ts_main_doubles(ts->side[0].nb, ts->chunk_max);
  const int m1 = ts->side[1].ncol_tot > 0 ? ts_main_doubles(ts->side[1].nb, ts->chunk_max) : 0;
  return (m0 + X_TOTAL + m1 + X_TOTAL) * 8;
}

int tb_launch_band_ts(const TsArgs& a, int smem, int num_sm, cudaStream_t st) {
  if (a.batch <= 0) return 0;
  int dev = 0;
  TB_CUDA(cudaGetDevice(&dev));
  auto kern = a.dim == 3 ? k_band_ts<3> : k_band_ts<2>;
  DevCache& dc = g_cache[dev & 63];
  if (dc.smem_set[a.dim - 2] < smem) {
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dc.smem_set[a.dim - 2] = smem;
  }
  if (num_sm <= 0) num_sm = 148;
  int per_sm = (228 * 1024) / (smem + 1024);
  if (per_sm > 7) per_sm = 7;
  if (per_sm < 1) per_sm = 1;
  int grid = a.batch < num_sm * per_sm ? a.batch : num_sm * per_sm;
  tb_prof_begin(TB_PROF_CHOL, st);
  kern<<<grid, 64, smem, st>>>(a);
  tb_prof_end(TB_PROF_CHOL, st);
  tb_count_launch(1);
  return (int)cudaGetLastError();
}
