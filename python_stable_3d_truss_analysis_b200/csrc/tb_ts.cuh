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
constexpr int32_t TB_Q_DUMMY = (int32_t)0x80000000u;   // tq_first of a hole in the entry schedule: negative (the assembly pass's
                                                         // first loop skips it) and absent from tq_multi (so does the second)
constexpr int TS_EPL = 5;      // K values per lane prefetched at the top of a block column (more: loaded on the spot)

struct TsSideDev {
  int ncol_own;              // block columns this side factorises on its own
  int ncol_tot;              // + separator columns (top: factorised after the hand-over; bottom: products only)
  int nb;                    // sub-diagonal blocks of this side's band view
  int ent0;                  // first entry of this side in program order (its block columns follow each other in kv / epos)
  const int4* colrec;        // [ncol_tot] one int4 per block column, all the kernel reads at the top of a column:
                             //              x = colmask | xmask << 9 | (number of K entries) << 18
                             //                  colmask: bit rb <=> block (c+rb, c) non-zero (bottom separator columns: blocks it contributes to)
                             //                  xmask:   top separator columns: blocks handed over by the bottom side
                             //              y = offset (doubles) of the factor chunk [Z^T | y | blocks]; its size follows from colmask
                             //              z, w = bit i <=> block product i of the column is structurally non-zero, i the flat index of
                             //              (distance d, block row rb) for NB = max(nb of both sides): d = 1..NB, rb = 0..NB-d
  const int32_t* rowdof;     // [ncol_tot*8] DOF index of virtual row v, -1 on padding
  const int32_t* rownat;     // [ncol_tot*8] internal (natural) row of virtual row v, -1 on padding
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
  const int32_t* epos;       // [nnz]          byte offset of the entry's staging position from the side's shared-memory base
  int64_t nnz;
};

struct TsSideHost {
  int ncol_own = 0, ncol_tot = 0, nb = 0;
  std::vector<uint32_t> colmask, srcmask, xmask;
  std::vector<int32_t> rowdof, rownat, lofs;
  std::vector<int4> colrec;
  std::vector<int2> colent;       // entries [x, y) per block column (program order)
  // device mirrors
  int32_t *d_rowdof = nullptr, *d_rownat = nullptr;
  int4* d_colrec = nullptr;
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
  std::vector<int32_t> epos;          // [nnz] staging position: byte offset from the side's shared-memory base (ts_stage_offset)
  std::vector<int32_t> ent_src;       // [nnz] index of the plan's scatter-map entry
  // assembly program in that order, in the format k_prep reads (tb_common.cuh: q_first / q_multi / q_ptr / q_pack)
  std::vector<int32_t> tq_first, tq_multi, tq_ptr, tq_pack;
  int32_t *d_epos = nullptr, *d_tq_first = nullptr, *d_tq_multi = nullptr, *d_tq_ptr = nullptr, *d_tq_pack = nullptr;
  int4* d_tq_multi4 = nullptr;        // {entry, first contribution, end, 0} of the entries with several contributions
};

// ---- shared-memory layout of one side (doubles): [main: ring of live blocks, later the back substitution's chunk buffers
// and u ring | 64 diagonal-block staging, then Z as a DMMA operand | 8 rhs | ring of the last y blocks | 16 misc]
constexpr int TS_NSTAGE = 3;                     // factor chunks in flight during the back substitution
constexpr int TS_X_SCR = 0, TS_X_T = TS_BE, TS_X_Y = TS_X_T + TS_BT, TS_X_MISC = TS_X_Y + (TS_NBX + 1) * TS_BT,
              TS_X_TOTAL = TS_X_MISC + 16;
__host__ __device__ inline int ts_main_doubles(int nb, int chunk_max) {
  const int ring = nb * (nb + 1) / 2 * TS_BE;
  const int back = TS_NSTAGE * chunk_max + (TS_NBX + 1) * TS_BT;   // (the u ring rotates through TS_NBX + 1 slots whatever nb is)
  return ring > back ? ring : back;
}
// Ring slot the new block (c+rb, c) of diagonal rb takes at block column c: the kernel's pointer ring starts with slot
// rb(rb-1)/2 + j - 1 for the block created j columns ago and always reuses the slot of the dying block, which makes the
// slot a function of c alone -- so the plan can resolve every K entry's staging address.
__host__ __device__ inline int ts_ring_slot(int rb, int c) { return rb * (rb - 1) / 2 + (rb - 1 - c % rb); }
// element (r, k) of an 8x8 block in operand-fragment layout: slab k / 4 holds [row 8][k 4]; the rows of slab 1 are
// permuted (r ^ 2) so that a warp's 16-byte stores of accumulator pairs (lane (qr, qc) -> elements (qr, 2qc), (qr, 2qc+1))
// fall into eight different 16-byte bank groups per quarter warp: slab 0 takes bytes [0, 64) of every 128, slab 1 [64, 128)
__host__ __device__ inline int ts_b8_off(int r, int k) { return ((k >> 2) << 5) + ((r ^ ((k >> 2) << 1)) << 2) + (k & 3); }

struct tb_plan;
int tb_ts_build(tb_plan* p);                  // host program (always), device mirrors when the plan has a device
void tb_ts_destroy(TsPlan* ts, bool device);
size_t tb_ts_workspace_bytes(const tb_plan* p, int batch);
void tb_ts_carve(TsArgs& t, const TsPlan* ts, void* ws, int total, int b0, double** kv);   // L | X | Z | uf | kv | status, for systems b0.. of a workspace of `total`
void tb_ts_fill_sides(TsArgs& t, const TsPlan* ts);
int tb_ts_smem_bytes(const TsPlan* ts);
int tb_launch_band_ts(const TsArgs& a, int smem, int num_sm, cudaStream_t st);
