// Host-side builder of the two-sided band program (tb_ts.cuh).  Pure integer work on the plan's scatter map
// (Truss.GetKMatrix's "+=" loop, slientruss3d/truss.py:307-316, restricted to the free DOFs of truss.py:343).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "tb_common.cuh"
#include "tb_ts.cuh"

namespace {

template <typename T>
int up(T** dptr, const std::vector<T>& h) {
  *dptr = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)dptr, bytes);
  if (e != cudaSuccess) return (int)e;
  if (!h.empty()) {
    e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

inline int popc(uint32_t x) { return __builtin_popcount(x); }
inline int topbit(uint32_t x) { return x ? 31 - __builtin_clz(x) : 0; }

struct Masks {
  bool valid = false;
  int nb[2] = {0, 0};
  std::vector<uint32_t> col[2], src[2], xm[2];
  double t_own[2] = {0, 0}, t_sep = 0, t_x = 0;      // forward time of the own columns, the separator columns, the hand-over
  double b_own[2] = {0, 0}, b_sep = 0;                // back substitution
  // end of the slower side: the top side runs T, S forward and S, T backward; the bottom side B forward, then (once the
  // separator displacements exist) B backward
  double total() const {
    const double top_end = t_own[0] + t_sep + b_sep + b_own[0];
    const double bot_end = std::max(t_own[1] + t_x, t_own[0] + t_sep + b_sep) + b_own[1];
    return std::max(top_end, bot_end);
  }
  int64_t products = 0, solves = 0;
};

// cost model of one block column (cycles of one warp with seven systems per SM, from the kernel's phase counters on B200):
// fixed part (staging, 8 pivots, stores), 8x8 block products (two DMMAs + operand loads), block solves; back substitution
constexpr double C_COL = 9700.0, C_PROD = 215.0, C_SOLVE = 300.0, C_BACK = 3100.0;

// Block masks of both sides for the split (bT, nS): kblk[bj] bit e <=> K_ff has an entry in block (bj+e, bj)
Masks make_masks(const std::vector<uint32_t>& kblk, int nblk, int bT, int nS) {
  Masks m;
  const int nB = nblk - bT - nS;
  const int tot0 = bT + nS, tot1 = nB > 0 ? nB + nS : 0;
  m.col[0].assign(tot0, 0); m.src[0].assign(tot0, 0); m.xm[0].assign(tot0, 0);
  m.col[1].assign(tot1, 0); m.src[1].assign(tot1, 0); m.xm[1].assign(tot1, 0);
  for (int bj = 0; bj < nblk; ++bj)
    for (int e = 0; e <= TS_NBX + 1 && bj + e < nblk; ++e) {
      if (!((kblk[bj] >> e) & 1u) && e != 0) continue;
      const int bi = bj + e;
      if (bi < tot0) {
        m.col[0][bj] |= 1u << e;
      } else {
        if (bj < bT) return m;                 // a row of B coupled to T: not a separator
        m.col[1][nblk - 1 - bi] |= 1u << e;    // virtual column of the bottom side, same block distance
      }
    }
  auto symbolic = [&](int s, int ncol_src, int tot) -> bool {
    for (int c = 0; c < tot; ++c) {
      const uint32_t mc = c < ncol_src ? m.col[s][c] : 0u;
      m.src[s][c] = mc;
      if (topbit(m.col[s][c]) > TS_NBX) return false;
      for (int e1 = 1; e1 <= TS_NBX; ++e1) {
        if (!((mc >> e1) & 1u)) continue;
        for (int e2 = e1; e2 <= TS_NBX; ++e2)
          if ((mc >> e2) & 1u) {
            if (c + e1 >= tot) return false;
            m.col[s][c + e1] |= 1u << (e2 - e1);
          }
      }
    }
    return true;
  };
  if (nB > 0) {
    for (int c = 0; c < nB; ++c) m.col[1][c] |= 1u;
    if (!symbolic(1, nB, tot1)) return m;
    // hand-over: virtual block (nB + j' + rb', nB + j') of the bottom side is the separator block (row J + rb', column J),
    // J = nS - 1 - j' - rb', of the top side
    for (int jq = 0; jq < nS; ++jq) {
      const uint32_t mx = m.col[1][nB + jq];
      for (int rb = 0; rb <= TS_NBX; ++rb)
        if ((mx >> rb) & 1u) {
          const int J = nS - 1 - jq - rb;
          if (J < 0) return m;
          m.xm[0][bT + J] |= 1u << rb;
          m.col[0][bT + J] |= 1u << rb;
        }
    }
  }
  for (int c = 0; c < tot0; ++c) m.col[0][c] |= 1u;
  if (!symbolic(0, tot0, tot0)) return m;
  for (int s = 0; s < 2; ++s) {
    int nb = 1;
    for (uint32_t x : m.col[s]) nb = std::max(nb, topbit(x));
    m.nb[s] = nb;
    if (nb > TS_NBX) return m;
  }
  // cost: products counted at the column that receives them
  for (int s = 0; s < 2; ++s) {
    const int tot = s == 0 ? tot0 : tot1, own = s == 0 ? bT : nB;
    std::vector<double> prod(tot, 0.0);
    for (int c = 0; c < tot; ++c) {
      const uint32_t mc = m.src[s][c];
      for (int e1 = 1; e1 <= TS_NBX; ++e1)
        if ((mc >> e1) & 1u)
          for (int e2 = e1; e2 <= TS_NBX; ++e2)
            if ((mc >> e2) & 1u) { prod[c + e1] += 1.0; m.products++; }
    }
    for (int c = 0; c < tot; ++c) {
      const int sol = std::max(0, popc(m.col[s][c] >> 2));
      if (c < own) {
        m.t_own[s] += C_COL + C_PROD * prod[c] + C_SOLVE * sol;
        m.b_own[s] += C_BACK;
        m.solves += popc(m.col[s][c] >> 2);
      } else if (s == 0) {
        m.t_sep += C_COL + C_PROD * prod[c] + C_SOLVE * sol;
        m.b_sep += C_BACK;
        m.solves += popc(m.col[s][c] >> 2);
      } else {
        m.t_x += 0.15 * C_COL + C_PROD * prod[c];
      }
    }
  }
  m.valid = true;
  return m;
}

}  // namespace

void tb_ts_destroy(TsPlan* ts, bool device) {
  if (!ts) return;
  if (device)
    for (int s = 0; s < 2; ++s) {
      TsSideHost& h = ts->side[s];
      cudaFree(h.d_colrec); cudaFree(h.d_rowdof); cudaFree(h.d_rownat);
    }
  if (device) {
    cudaFree(ts->d_tq_multi4);
    cudaFree(ts->d_epos); cudaFree(ts->d_tq_first); cudaFree(ts->d_tq_multi); cudaFree(ts->d_tq_ptr); cudaFree(ts->d_tq_pack);
  }
  delete ts;
}

int tb_ts_build(tb_plan* p) {
  TsPlan* ts = new TsPlan();
  p->ts = ts;
  const int n = p->n, d = p->dim;
  if (n <= 0 || p->M <= 0) return 0;
  const size_t nnz = p->int_row.size();
  const int nblk = (n + TS_BT - 1) / TS_BT;
  ts->nblk = nblk;
  ts->n_pad = nblk * TS_BT;
  // block structure of K_ff and the reach of every block row (first block column of its envelope)
  std::vector<uint32_t> kblk(nblk, 0u);
  for (size_t e = 0; e < nnz; ++e) {
    const int bi = p->int_row[e] / TS_BT, bj = p->int_col[e] / TS_BT;
    if (bi - bj > TS_NBX) return 0;              // band too wide for this kernel: the other pipelines take it
    kblk[bj] |= 1u << (bi - bj);
  }
  std::vector<int> reach(nblk);                  // lowest block column touched by block row i
  std::iota(reach.begin(), reach.end(), 0);
  for (int bj = 0; bj < nblk; ++bj)
    for (int e = 0; e <= TS_NBX && bj + e < nblk; ++e)
      if ((kblk[bj] >> e) & 1u) reach[bj + e] = std::min(reach[bj + e], bj);

  // ---- choose the split: one-sided (everything on the top side), or the two-sided split with the smallest critical path
  Masks best = make_masks(kblk, nblk, nblk, 0);
  if (!best.valid) return 0;
  int best_bT = nblk, best_nS = 0;
  const double t_one = best.total();
  double best_t = t_one;
  const char* env1 = getenv("TB_TS_ONE_SIDED");
  const bool allow_two = !(env1 && env1[0] == '1');
  const char* envs = getenv("TB_TS_SPLIT");       // force the first separator block (experiments)
  const int force_bT = envs ? atoi(envs) : -1;
  if (allow_two)
    for (int bT = 1; bT + 2 <= nblk; ++bT) {
      if (force_bT >= 0 && bT != force_bT) continue;
      int last = bT - 1;                           // last block row whose envelope reaches into T
      for (int bi = bT; bi < nblk; ++bi)
        if (reach[bi] < bT) last = bi;
      const int nS = last - bT + 1;
      if (nS < 1 || nS > TS_NBX || bT + nS >= nblk) continue;
      Masks m = make_masks(kblk, nblk, bT, nS);
      if (!m.valid) continue;
      const double t = m.total();
      if (t < (force_bT >= 0 ? 1e300 : 0.9 * t_one) && (t < best_t || best_nS == 0)) {
        best = m; best_bT = bT; best_nS = nS; best_t = t;
      }
    }
  const int bT = best_bT, nS = best_nS, nB = nblk - bT - nS;
  ts->bT = bT; ts->nS = nS; ts->nB = nB;
  ts->products = best.products;
  ts->solves = best.solves;
  // executed tensor work: two DMMA (512 flop each) per block product and per block solve
  ts->dmma_flops = 1024.0 * (double)(best.products + best.solves);

  const int n_pad = ts->n_pad;
  for (int s = 0; s < 2; ++s) {
    TsSideHost& h = ts->side[s];
    h.ncol_own = s == 0 ? bT : nB;
    h.ncol_tot = s == 0 ? bT + nS : (nB > 0 ? nB + nS : 0);
    h.nb = best.nb[s];
    h.colmask = best.col[s];
    h.srcmask = best.src[s];
    h.xmask = best.xm[s];
    h.rowdof.assign((size_t)h.ncol_tot * TS_BT, -1);
    h.rownat.assign((size_t)h.ncol_tot * TS_BT, -1);
    for (int v = 0; v < h.ncol_tot * TS_BT; ++v) {
      const int nat = s == 0 ? v : n_pad - 1 - v;
      if (nat >= 0 && nat < n) {
        h.rownat[v] = nat;
        h.rowdof[v] = p->free_int[nat];
      }
    }
  }
  // ---- factor storage: top columns (separator included), then the bottom's own columns
  {
    int64_t off = 0;
    int cmax = 0;
    for (int s = 0; s < 2; ++s) {
      TsSideHost& h = ts->side[s];
      const int ncol = s == 0 ? h.ncol_tot : h.ncol_own;
      h.lofs.assign((size_t)h.ncol_tot + 1, 0);
      for (int c = 0; c < h.ncol_tot; ++c) {
        h.lofs[c] = (int32_t)off;
        if (c < ncol) {
          const int sz = TS_BE + TS_BT + TS_BE * popc(h.colmask[c] >> 1);
          off += sz;
          cmax = std::max(cmax, sz);
        }
      }
      h.lofs[h.ncol_tot] = (int32_t)off;
    }
    ts->l_per_sys = off;
    ts->chunk_max = cmax;
  }
  // ---- entries and members per side / block column / chunk
  struct Ent { int vr, vc; int32_t src; };
  std::vector<std::vector<Ent>> colent[2];
  colent[0].resize(ts->side[0].ncol_tot);
  colent[1].resize(ts->side[1].ncol_tot);
  for (size_t e = 0; e < nnz; ++e) {
    const int i = p->int_row[e], j = p->int_col[e];
    if (i / TS_BT < bT + nS) {
      colent[0][j / TS_BT].push_back({i, j, (int32_t)e});
    } else {
      const int vr = n_pad - 1 - j, vc = n_pad - 1 - i;
      colent[1][vc / TS_BT].push_back({vr, vc, (int32_t)e});
    }
  }
  ts->epos.clear(); ts->ent_src.clear();
  ts->tq_ptr.assign(1, 0);
  for (int s = 0; s < 2; ++s) {
    TsSideHost& h = ts->side[s];
    h.colrec.resize((size_t)h.ncol_tot);
    h.colent.resize(h.ncol_tot);
    const int NBK = std::max(ts->side[0].nb, ts->side[1].nb);   // the kernel instantiation (its flat product list)
    const int mainsz = ts_main_doubles(h.nb, ts->chunk_max);
    for (int c = 0; c < h.ncol_tot; ++c) {
      const int e0 = (int)ts->epos.size();
      // Entry order inside the block column: the kernel adds entry e0 + 32 i + lane in round i with one 8-byte shared-memory
      // read-modify-write per lane, which the hardware serves per half warp.  Entries are dealt out so that the 16 lanes of
      // a half warp hit different 8-byte bank pairs (address / 8 mod 16) wherever the bucket sizes allow it; holes are dummy
      // entries (epos -1, no contributions, never read).
      std::vector<int32_t> pos(colent[s][c].size());
      std::vector<std::vector<int>> bucket(16);
      for (size_t i = 0; i < colent[s][c].size(); ++i) {
        const Ent& en = colent[s][c][i];
        const int rb = en.vr / TS_BT - en.vc / TS_BT;
        const int slot_off = rb == 0 ? mainsz + TS_X_SCR : ts_ring_slot(rb, c) * TS_BE;
        pos[i] = (slot_off + ts_b8_off(en.vr % TS_BT, en.vc % TS_BT)) * 8;
        bucket[(pos[i] >> 3) & 15].push_back((int)i);
      }
      // the fewest rounds the entries need; inside them, every bucket is spread over the half-rounds as evenly as possible
      // (largest buckets first, each entry to the half-round holding the fewest entries of its bucket, then the fewest at all)
      size_t halves = 2 * ((colent[s][c].size() + 31) / 32);
      std::vector<std::vector<int>> half(halves);
      // Measured on B200 (bar-942 x 1024): the bank-aware order takes 3 us off the band kernel (0.256 -> 0.253 ms) and adds
      // 18 us to the assembly pass (0.064 -> 0.082 ms: its shared-memory reads of the member products lose the locality of
      // the scatter map's order), so the scatter map's order is the default; TB_TS_BANK_ORDER=1 selects the other one.
      static const bool bank_order = getenv("TB_TS_BANK_ORDER") && getenv("TB_TS_BANK_ORDER")[0] == '1';
      if (!bank_order) {
        for (size_t i = 0; i < colent[s][c].size(); ++i) half[i / 16].push_back((int)i);
      } else {
        std::vector<int> order(16);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bucket[a].size() > bucket[b].size(); });
        std::vector<int> mult(halves);
        for (int k : order) {
          std::fill(mult.begin(), mult.end(), 0);
          for (int i : bucket[k]) {
            size_t bestp = halves;
            for (size_t hr = 0; hr < halves; ++hr) {
              if (half[hr].size() >= 16) continue;
              if (bestp == halves || mult[hr] < mult[bestp] || (mult[hr] == mult[bestp] && half[hr].size() < half[bestp].size())) bestp = hr;
            }
            half[bestp].push_back(i);
            mult[bestp]++;
          }
        }
      }
      for (size_t hr = 0; hr < halves; ++hr)
        for (int k = 0; k < 16; ++k) {
          const int i = k < (int)half[hr].size() ? half[hr][k] : -1;
          if (i < 0) {
            ts->epos.push_back(-1);
            ts->ent_src.push_back(-1);
          } else {
            const Ent& en = colent[s][c][i];
            ts->epos.push_back(pos[i]);
            ts->ent_src.push_back(en.src);
            for (int64_t k2 = p->ent_ptr[en.src]; k2 < p->ent_ptr[en.src + 1]; ++k2) {
              const int loc = p->ctr_local[k2], la = loc / (2 * d), lb = loc % (2 * d);
              const int A = la / d, i2 = la % d, B = lb / d, j2 = lb % d;
              const int lo = std::min(i2, j2), hi = std::max(i2, j2);
              ts->tq_pack.push_back((p->ctr_member[k2] << 4) | ((A != B) << 3) | (lo * d - lo * (lo - 1) / 2 + (hi - lo)));
            }
          }
          ts->tq_ptr.push_back((int32_t)ts->tq_pack.size());
        }
      while (!ts->epos.empty() && (int)ts->epos.size() > e0 && ts->epos.back() < 0) {   // trailing holes of the last half-round
        ts->epos.pop_back(); ts->ent_src.pop_back(); ts->tq_ptr.pop_back();
      }
      h.colent[c] = make_int2(e0, (int)ts->epos.size());
      // block products of this column: L(c+rb, c-d) L(c, c-d)^T needs both blocks of column c-d
      uint64_t pm = 0;
      int idx = 0;
      for (int dd = 1; dd <= NBK; ++dd)
        for (int rb = 0; rb + dd <= NBK; ++rb, ++idx) {
          if (c - dd < 0) continue;
          const uint32_t sm = h.srcmask[c - dd];
          if (((sm >> dd) & 1u) && ((sm >> (rb + dd)) & 1u)) pm |= (uint64_t)1 << idx;
        }
      h.colrec[c] = make_int4((int)(h.colmask[c] | (h.xmask[c] << 9) | ((uint32_t)((int)ts->epos.size() - e0) << 18)), h.lofs[c],
                              (int)(uint32_t)pm, (int)(uint32_t)(pm >> 32));
    }
  }
  {   // first contribution inline, entries with several contributions listed separately (as tb_plan.cu does for the other orders)
    const size_t nq = ts->epos.size();
    ts->tq_first.assign(nq, 0);
    ts->tq_multi.clear();
    for (size_t q = 0; q < nq; ++q) {
      const int cnt = ts->tq_ptr[q + 1] - ts->tq_ptr[q];
      ts->tq_first[q] = cnt > 0 ? (ts->tq_pack[ts->tq_ptr[q]] | (cnt > 1 ? (int32_t)0x80000000u : 0)) : TB_Q_DUMMY;
      if (cnt > 1) ts->tq_multi.push_back((int32_t)q);
    }
  }
  ts->ok = 1;
  if (p->device < 0) return 0;
  int rc = 0;
  for (int s = 0; s < 2 && !rc; ++s) {
    TsSideHost& h = ts->side[s];
    if (!rc) rc = up(&h.d_colrec, h.colrec);
    if (!rc) rc = up(&h.d_rowdof, h.rowdof);
    if (!rc) rc = up(&h.d_rownat, h.rownat);
  }
  if (!rc) rc = up(&ts->d_epos, ts->epos);
  if (!rc) rc = up(&ts->d_tq_first, ts->tq_first);
  if (!rc) rc = up(&ts->d_tq_multi, ts->tq_multi);
  if (!rc) rc = up(&ts->d_tq_ptr, ts->tq_ptr);
  if (!rc) rc = up(&ts->d_tq_pack, ts->tq_pack);
  if (!rc) {
    std::vector<int4> m4(ts->tq_multi.size());
    for (size_t i = 0; i < m4.size(); ++i) {
      const int q = ts->tq_multi[i];
      m4[i] = make_int4(q, ts->tq_ptr[q], ts->tq_ptr[q + 1], 0);
    }
    rc = up(&ts->d_tq_multi4, m4);
  }
  return rc;
}

size_t tb_ts_workspace_bytes(const tb_plan* p, int batch) {
  const TsPlan* ts = p->ts;
  if (!ts || !ts->ok) return 0;
  const size_t per = (size_t)ts->l_per_sys + (size_t)ts->nS * ts->nS * TS_BE + (size_t)ts->nS * TS_BT + (size_t)ts->n_pad +
                     ts->epos.size();
  return (size_t)batch * per * 8 + (size_t)batch * 4 + 4096;
}

void tb_ts_carve(TsArgs& t, const TsPlan* ts, void* ws, int total, int b0, double** kv) {
  double* p = (double*)ws;
  t.L = p + (size_t)b0 * ts->l_per_sys;               p += (size_t)total * ts->l_per_sys;
  t.X = p + (size_t)b0 * ts->nS * ts->nS * TS_BE;     p += (size_t)total * ts->nS * ts->nS * TS_BE;
  t.Z = p + (size_t)b0 * ts->nS * TS_BT;              p += (size_t)total * ts->nS * TS_BT;
  t.uf = p + (size_t)b0 * ts->n_pad;                  p += (size_t)total * ts->n_pad;
  *kv = p + (size_t)b0 * ts->epos.size();             p += (size_t)total * ts->epos.size();
  t.kv = *kv;
  t.status = (int32_t*)p + b0;
}

void tb_ts_fill_sides(TsArgs& t, const TsPlan* ts) {
  for (int s = 0; s < 2; ++s) {
    const TsSideHost& h = ts->side[s];
    TsSideDev& d = t.side[s];
    d.ncol_own = h.ncol_own; d.ncol_tot = h.ncol_tot; d.nb = h.nb;
    d.colrec = h.d_colrec; d.rowdof = h.d_rowdof; d.rownat = h.d_rownat;
    d.ent0 = h.colent.empty() ? 0 : h.colent[0].x;
  }
  t.epos = ts->d_epos;
  t.nnz = (int64_t)ts->epos.size();
  t.nS = ts->nS;
  t.chunk_max = ts->chunk_max;
  t.l_per_sys = ts->l_per_sys;
  t.n_pad = ts->n_pad;
}

// ---- debug / test exports: the program as flat host arrays (tests/test_ts_program_cpu.py replays it in numpy)
extern "C" int tb_plan_ts_info(const tb_plan* p, int32_t* out /*[16]*/) {
  if (!p || !out) return TB_ERR_NULL;
  const TsPlan* ts = p->ts;
  for (int i = 0; i < 16; ++i) out[i] = 0;
  if (!ts || !ts->ok) return TB_OK;
  out[0] = 1; out[1] = ts->nblk; out[2] = ts->n_pad; out[3] = ts->bT; out[4] = ts->nS; out[5] = ts->nB;
  out[6] = ts->side[0].nb; out[7] = ts->side[1].nb; out[8] = ts->chunk_max; out[9] = (int32_t)ts->l_per_sys;
  out[10] = (int32_t)ts->products; out[11] = (int32_t)ts->solves;
  out[12] = (int32_t)ts->epos.size();
  out[13] = (int32_t)ts->tq_pack.size();
  out[14] = (int32_t)ts->tq_multi.size();
  return TB_OK;
}

// which = 0 colmask, 1 srcmask, 2 xmask, 3 colent (2 ints each), 4 rowdof, 5 rownat, 6 lofs (per side); 7 epos, 8 ent_src,
// 9 tq_first, 10 tq_multi, 11 tq_ptr, 12 tq_pack (whole program, side ignored); 13 colrec (4 ints per block column, per side).  Returns the element count (ints); copies
// when out != NULL.
extern "C" int64_t tb_plan_ts_array(const tb_plan* p, int32_t side, int32_t which, int32_t* out) {
  if (!p || !p->ts || !p->ts->ok || side < 0 || side > 1) return -1;
  const TsPlan* ts = p->ts;
  const TsSideHost& h = ts->side[side];
  const void* src = nullptr;
  size_t cnt = 0;
  switch (which) {
    case 0: src = h.colmask.data(); cnt = h.colmask.size(); break;
    case 1: src = h.srcmask.data(); cnt = h.srcmask.size(); break;
    case 2: src = h.xmask.data(); cnt = h.xmask.size(); break;
    case 3: src = h.colent.data(); cnt = h.colent.size() * 2; break;
    case 4: src = h.rowdof.data(); cnt = h.rowdof.size(); break;
    case 5: src = h.rownat.data(); cnt = h.rownat.size(); break;
    case 6: src = h.lofs.data(); cnt = h.lofs.size(); break;
    case 7: src = ts->epos.data(); cnt = ts->epos.size(); break;
    case 8: src = ts->ent_src.data(); cnt = ts->ent_src.size(); break;
    case 9: src = ts->tq_first.data(); cnt = ts->tq_first.size(); break;
    case 10: src = ts->tq_multi.data(); cnt = ts->tq_multi.size(); break;
    case 11: src = ts->tq_ptr.data(); cnt = ts->tq_ptr.size(); break;
    case 12: src = ts->tq_pack.data(); cnt = ts->tq_pack.size(); break;
    case 13: src = h.colrec.data(); cnt = h.colrec.size() * 4; break;
    default: return -1;
  }
  if (out && cnt) memcpy(out, src, cnt * 4);
  return (int64_t)cnt;
}
