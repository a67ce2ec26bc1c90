// Two-sided band program (tb_tsplan.cu builds it on the host, tb_bandts.cu runs it).
//
// The reduced stiffness matrix K_ff (internal elimination order, half-bandwidth <= 8 blocks of 8) is split into
//   T  = block rows [0, bT)            eliminated top-down by warp 0            (side 0, "top")
//   S  = block rows [bT, bT + nS)      the separator: every row coupled to T that is not in T
//   B  = block rows [bT + nS, nblk)    eliminated bottom-up by warp 1           (side 1, "bottom")
// which is the nested-dissection order [T, reverse(B), S] of a quasi one-dimensional structure: the two chains are
// independent until the separator, so the dependent pivot chain is max(|T|, |B|) + |S| instead of n.  Each side sees an
// ordinary band matrix in its own VIRTUAL numbering (top: v = row; bottom: v = n_pad - 1 - row, which turns the
// bottom-up elimination into a plain Cholesky), so one column routine serves both.  The top side goes on through the
// separator columns; the bottom side only forms its Schur-complement contribution to them (products, no factor) and hands
// it over through global scratch in the top side's orientation.
//
// Per side and virtual block column c the program lists
//   colmask   bit rb: block (c+rb, c) of L is structurally non-zero (block symbolic factorisation of this order)
//   entries   the K_ff entries of the block column: a range [e0, e1) of the "program order" of all entries, with
//             epos[e] = block offset rb << 6 | position inside the 8x8 block.  The K values arrive in the same order
//             (kv[e]) from the assembly pass (tb_large.cu k_prep run on tq_*: every entry sums its member contributions
//             in ascending member order, truss.py:310-314), so a lane's loads are coalesced and nothing is searched
//   lofs      where the column's chunk of the factor lives ([Z = L_D^{-T} | y | non-zero blocks below the diagonal])
// An earlier version assembled K inside the band kernel (member geometry + gather per block column); that was 35 % of
// the kernel's instructions on its dependent chain, so the assembly went back to its own (parallel, HBM-bound) pass.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

constexpr int TS_BT = 8;       // block order: the native FP64 MMA shape (mma.m8n8k4)
constexpr int TS_BE = 64;      // doubles per block
constexpr int TS_NBX = 8;      // at most this many sub-diagonal blocks per block column
constexpr int TS_EPL = 5;      // K values per lane prefetched at the top of a block column (more: loaded on the spot)

struct TsSideDev {
  int ncol_own;              // block columns this side factorises on its own
  int ncol_tot;              // + separator columns (top: factorised after the hand-over; bottom: products only)
  int nb;                    // sub-diagonal blocks of this side's band view
  const int4* colinfo;       // [ncol_tot]   x = colmask: bit rb <=> block (c+rb, c) non-zero (bottom separator columns: blocks it contributes to)
                             //              y = srcmask: colmask of the columns that exist as factor columns (0 for the bottom's separator columns)
                             //              z = xmask: top separator columns: blocks handed over by the bottom side;  w = unused
  const int2* colent;        // [ncol_tot]   entries [x, y) of the block column in program order (indices into kv / epos)
  const int32_t* rowdof;     // [ncol_tot*8] DOF index of virtual row v, -1 on padding
  const int32_t* rownat;     // [ncol_tot*8] internal (natural) row of virtual row v, -1 on padding
  const int32_t* lofs;       // [ncol_tot+1] offset (doubles) of the column's factor chunk inside the system's factor storage
};

struct TsArgs {
  int batch, dim, nJ, M, N, n, n_pad;
  const double* xyz;      int64_t xyz_stride;
  const double* aed;      int64_t aed_stride;
  const int32_t* gene;    int64_t gene_stride;
  const double* type_table; int n_type;
  const double* force;    int64_t force_stride;
  TsSideDev side[2];
  int nS;                    // separator block columns
  int chunk_max;             // largest factor chunk (doubles)
  int64_t l_per_sys;         // factor storage per system (doubles)
  double* L;                 // [B][l_per_sys]
  double* X;                 // [B][nS*nS*64]  bottom side's Schur contribution to the separator, top orientation
  double* Z;                 // [B][nS*8]      ... and to the forward-substituted right-hand side
  double* uf;                // [B][n_pad]     free displacements, internal order (read by the recovery)
  int32_t* status;           // [B]
  const double* kv;          // [B][nnz]       K_ff values in program order (assembly pass)
  const int32_t* epos;       // [nnz]          rb << 6 | position inside the block (operand-fragment layout)
  int64_t nnz;
};

struct TsSideHost {
  int ncol_own = 0, ncol_tot = 0, nb = 0;
  std::vector<uint32_t> colmask, srcmask, xmask;
  std::vector<int32_t> rowdof, rownat, lofs;
  std::vector<int4> colinfo;
  std::vector<int2> colent;
  // device mirrors
  int32_t *d_rowdof = nullptr, *d_rownat = nullptr, *d_lofs = nullptr;
  int4* d_colinfo = nullptr;
  int2* d_colent = nullptr;
};

struct TsPlan {
  int ok = 0;               // the program exists (band narrow enough, splits found)
  int nblk = 0, n_pad = 0;
  int bT = 0, nS = 0, nB = 0;
  int chunk_max = 0;
  int64_t l_per_sys = 0;
  int64_t products = 0, solves = 0;   // 8x8 block products / block solves per system (executed DMMA work)
  double dmma_flops = 0.0;            // flops the tensor cores execute per system
  TsSideHost side[2];
  // entries of both sides in program order (top side's block columns, then the bottom side's)
  std::vector<int32_t> epos;          // [nnz] rb << 6 | position inside the block
  std::vector<int32_t> ent_src;       // [nnz] index of the plan's scatter-map entry
  // assembly program in that order, in the format k_prep reads (tb_common.cuh: q_first / q_multi / q_ptr / q_pack)
  std::vector<int32_t> tq_first, tq_multi, tq_ptr, tq_pack;
  int32_t *d_epos = nullptr, *d_tq_first = nullptr, *d_tq_multi = nullptr, *d_tq_ptr = nullptr, *d_tq_pack = nullptr;
};

struct tb_plan;
int tb_ts_build(tb_plan* p);                  // host program (always), device mirrors when the plan has a device
void tb_ts_destroy(TsPlan* ts, bool device);
size_t tb_ts_workspace_bytes(const tb_plan* p, int batch);
void tb_ts_carve(TsArgs& t, const TsPlan* ts, void* ws, int batch, double** kv);   // L | X | Z | uf | kv | status
void tb_ts_fill_sides(TsArgs& t, const TsPlan* ts);
int tb_ts_smem_bytes(const TsPlan* ts);
int tb_launch_band_ts(const TsArgs& a, int smem, int num_sm, cudaStream_t st);
