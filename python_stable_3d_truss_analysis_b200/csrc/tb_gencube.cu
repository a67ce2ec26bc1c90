// Random cube-truss topologies on the device: the generator of slientruss3d/generate.py
//   CubeGrid.RandomGenerateCubes :266-287   random walk over a grid of unit cells (DFS / BFS / random end of the frontier)
//   CubeTruss.GenerateNew        :176-185   joint ids in first-seen order over the cubes' corners
//   CubeTruss.LinkMember         :187-232   one or both diagonals per face, the twelve edges, duplicates dropped
//   CubeGrid.ProcessPinSupport   :289-300   positions = corner * cell length, lowest layer pinned
//   AssignRandomForces / AssignRandomMemberType :318-336, the stability re-draw :343-372
// as one thread per truss.  The walk is sequential by nature but tiny (a few tens of steps on bitmasks), so the
// parallelism is across trusses.  Random numbers are counter-based (Philox4x32-10 keyed by the seed; counter = truss,
// attempt, draw), so a dataset is reproducible from (seed, index) but does NOT replay Python's `random` stream -- the
// host classes in generate.py keep that property and are the oracle of this kernel's deterministic parts: the kernel
// exports the cell sequence and the diagonal picks of every truss, and the tests replay them through CubeTruss /
// CubeGrid.CubesToTruss and compare joints, supports and members exactly.
// Every truss is written at a fixed stride (max_joint, max_member); tb_gencube_pack compacts the batch into the packed
// layout of tb_ragged_in once the per-truss counts have been prefix-summed.
#include <math.h>

#include "tb_common.cuh"

namespace {

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

constexpr uint32_t PUR_GEN = 21;

// sequential stream of one (truss, attempt): draw i is word i % 4 of Philox block i / 4
struct Stream {
  uint32_t o, attempt, n = 0;
  uint64_t seed;
  U4 buf;
  __device__ uint32_t next() {
    if ((n & 3u) == 0) buf = philox(o, attempt, n >> 2, PUR_GEN, seed);
    const uint32_t v = (n & 3u) == 0 ? buf.x : (n & 3u) == 1 ? buf.y : (n & 3u) == 2 ? buf.z : buf.w;
    ++n;
    return v;
  }
  __device__ int below(int m) { return (int)(((uint64_t)next() * (uint64_t)m) >> 32); }   // uniform in [0, m)
  __device__ double u01() {
    const uint32_t a = next(), b = next();
    return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * u01(); }       // random.uniform
};

constexpr int MAXC = TB_GEN_MAX_CELLS, MAXV = TB_GEN_MAX_VERTS;

// corner pairs of LinkMember: six faces x two diagonals, then the twelve edges (generate.py:213-231)
__constant__ int8_t c_diag[6][2][2] = {{{0, 5}, {1, 4}}, {{1, 7}, {3, 5}}, {{3, 6}, {2, 7}}, {{2, 4}, {0, 6}}, {{4, 7}, {5, 6}}, {{0, 3}, {1, 2}}};
__constant__ int8_t c_edge[12][2] = {{4, 5}, {5, 7}, {6, 7}, {4, 6}, {0, 1}, {0, 2}, {1, 3}, {2, 3}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

struct GenArgs {
  tb_gencube_params p;
  int n, max_joint, max_member, max_cube;
  const double* type_table;
  double* xyz; uint8_t* support; double* force; int32_t* conn; double* aed;
  int32_t* n_joint; int32_t* n_member; int32_t* info;
  int16_t* cells; uint8_t* picks; double* length;          // optional exports: [n][max_cube], [n][max_cube][6], [n][3]
};

__global__ void __launch_bounds__(64) k_gencube(const GenArgs a) {
  const tb_gencube_params& P = a.p;
  const int gx = P.grid[0], gy = P.grid[1], gz = P.grid[2];
  const int ncell = gx * gy * gz, vx = gx + 1, vy = gy + 1;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < a.n; o += gridDim.x * blockDim.x) {
    double* xyz = a.xyz + (int64_t)o * a.max_joint * 3;
    uint8_t* sup = a.support + (int64_t)o * a.max_joint;
    double* force = a.force + (int64_t)o * a.max_joint * 3;
    int32_t* conn = a.conn + (int64_t)o * a.max_member * 2;
    double* aed = a.aed + (int64_t)o * a.max_member * 3;
    int16_t* cells = a.cells ? a.cells + (int64_t)o * a.max_cube : nullptr;
    uint8_t* picks = a.picks ? a.picks + (int64_t)o * a.max_cube * 6 : nullptr;
    int nJ = 0, M = 0, status = TB_INFO_NOT_STABLE;
    for (int attempt = 0; attempt < P.max_attempts; ++attempt) {
      Stream rng;
      rng.o = (uint32_t)o; rng.attempt = (uint32_t)attempt; rng.seed = P.seed;
      uint32_t used[(MAXC + 31) / 32] = {}, infr[(MAXC + 31) / 32] = {};
      uint32_t linked[(MAXV * 9 + 31) / 32] = {};
      int16_t frontier[MAXC];
      int16_t vid[MAXV];
      for (int i = 0; i < MAXV; ++i) vid[i] = -1;
      const int want = P.ncube_lo + rng.below(P.ncube_hi - P.ncube_lo + 1);
      // ---- RandomGenerateCubes: the frontier is an ordered list (pop from either end, extend at the back)
      int fhead = 0, ftail = 0;                 // live entries: frontier[fhead .. ftail)
      frontier[ftail++] = (int16_t)rng.below(ncell);        // GetRandomFeasible on an empty grid
      infr[frontier[0] >> 5] |= 1u << (frontier[0] & 31);
      int ncube = 0;
      nJ = 0; M = 0;
      int minz = 1 << 30;
      // (joints are numbered and members linked as each cube is created: LinkMember only needs the joint ids of its own
      // cube and the set of links made so far, so CubesToTruss is fused into the walk)
      while (ncube < want && fhead < ftail) {
        int cell;
        const bool back = P.method == 0 ? true : P.method == 1 ? false : (rng.u01() <= 0.5);
        if (back) cell = frontier[--ftail];
        else cell = frontier[fhead++];
        infr[cell >> 5] &= ~(1u << (cell & 31));
        used[cell >> 5] |= 1u << (cell & 31);
        const int cx = cell % gx, cy = (cell / gx) % gy, cz = cell / (gx * gy);
        // GetNextFeasibles: -x, +x, -y, +y, -z, +z, shuffled (Fisher-Yates), those not yet in the frontier appended
        int16_t nb[6];
        int nn = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax)
#pragma unroll
          for (int st = -1; st <= 1; st += 2) {
            const int x = cx + (ax == 0 ? st : 0), y = cy + (ax == 1 ? st : 0), z = cz + (ax == 2 ? st : 0);
            if (x < 0 || x >= gx || y < 0 || y >= gy || z < 0 || z >= gz) continue;
            const int c2 = x + gx * (y + gy * z);
            if ((used[c2 >> 5] >> (c2 & 31)) & 1u) continue;
            nb[nn++] = (int16_t)c2;
          }
        for (int i = nn - 1; i > 0; --i) {
          const int j = rng.below(i + 1);
          const int16_t t = nb[i]; nb[i] = nb[j]; nb[j] = t;
        }
        // (a cell enters the frontier at most once, so the list never outgrows the array)
        for (int i = 0; i < nn; ++i) {
          const int c2 = nb[i];
          if ((infr[c2 >> 5] >> (c2 & 31)) & 1u) continue;
          infr[c2 >> 5] |= 1u << (c2 & 31);
          frontier[ftail++] = (int16_t)c2;
        }
        // ---- CubeTruss.GenerateNew: corner k = cell + ((k >> axis) & 1)
        int jid[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int x = cx + (k & 1), y = cy + ((k >> 1) & 1), z = cz + ((k >> 2) & 1);
          const int v = x + vx * (y + vy * z);
          if (vid[v] < 0) {
            vid[v] = (int16_t)nJ;
            xyz[3 * nJ] = (double)x; xyz[3 * nJ + 1] = (double)y; xyz[3 * nJ + 2] = (double)z;   // grid units for now
            ++nJ;
          }
          jid[k] = vid[v];
          if (z < minz) minz = z;
        }
        // ---- LinkMember
        auto link = [&](int ka, int kb) {
          if (!P.allow_parallel) {
            const int xa = cx + (ka & 1), ya = cy + ((ka >> 1) & 1), za = cz + ((ka >> 2) & 1);
            const int dx = (kb & 1) - (ka & 1), dy = ((kb >> 1) & 1) - ((ka >> 1) & 1), dz = ((kb >> 2) & 1) - ((ka >> 2) & 1);
            // the nine link directions: +x +y +z | (+x,+z) (-x,+z) (+y,+z) (-y,+z) (+x,+y) (-x,+y)
            int type;
            if (dy == 0 && dz == 0) type = 0;
            else if (dx == 0 && dz == 0) type = 1;
            else if (dx == 0 && dy == 0) type = 2;
            else if (dy == 0) type = dx > 0 ? 3 : 4;
            else if (dx == 0) type = dy > 0 ? 5 : 6;
            else type = dx > 0 ? 7 : 8;
            const int key = (xa + vx * (ya + vy * za)) * 9 + type;
            if ((linked[key >> 5] >> (key & 31)) & 1u) return;
            linked[key >> 5] |= 1u << (key & 31);
          }
          conn[2 * M] = jid[ka];
          conn[2 * M + 1] = jid[kb];
          ++M;
        };
        for (int f = 0; f < 6; ++f) {
          const int pick = P.link_type == 3 ? rng.below(3) : P.link_type;
          if (picks) picks[ncube * 6 + f] = (uint8_t)pick;
          if (pick == 0 || pick == 2) link(c_diag[f][0][0], c_diag[f][0][1]);
          if (pick == 1 || pick == 2) link(c_diag[f][1][0], c_diag[f][1][1]);
        }
        for (int e = 0; e < 12; ++e) link(c_edge[e][0], c_edge[e][1]);
        if (cells) cells[ncube] = (int16_t)cell;
        ++ncube;
      }
      if (cells)
        for (int i = ncube; i < a.max_cube; ++i) cells[i] = -1;
      // ---- ProcessPinSupport: cell lengths, lowest layer pinned and moved to z = 0
      double L[3];
      for (int i = 0; i < 3; ++i) L[i] = rng.uniform(P.length_lo, P.length_hi);
      if (a.length)
        for (int i = 0; i < 3; ++i) a.length[(int64_t)o * 3 + i] = L[i];
      int npin = 0;
      for (int j = 0; j < nJ; ++j) {
        const double z = xyz[3 * j + 2];
        const bool pin = P.add_pin && (int)z == minz;
        sup[j] = pin ? 1 : 0;
        npin += pin;
        xyz[3 * j] = xyz[3 * j] * L[0];
        xyz[3 * j + 1] = xyz[3 * j + 1] * L[1];
        xyz[3 * j + 2] = (z - (double)minz) * L[2];
        force[3 * j] = force[3 * j + 1] = force[3 * j + 2] = 0.0;
      }
      // ---- AssignRandomForces: nForce of the unsupported joints, a uniform vector each
      const int nfree = nJ - npin;
      if (nfree > 0) {
        int lo = P.nforce_lo < 0 ? 1 : P.nforce_lo, hi = P.nforce_hi < 0 ? nfree : P.nforce_hi;
        if (hi > nfree) hi = nfree;
        if (lo > hi) lo = hi;
        const int nforce = lo + rng.below(hi - lo + 1);
        // partial Fisher-Yates over the unsupported joints (vid is free now: reuse it as the index list)
        int cnt = 0;
        for (int j = 0; j < nJ; ++j)
          if (!sup[j]) vid[cnt++] = (int16_t)j;
        for (int i = 0; i < nforce; ++i) {
          const int j = i + rng.below(cnt - i);
          const int16_t t = vid[i]; vid[i] = vid[j]; vid[j] = t;
          for (int ax = 0; ax < 3; ++ax) force[3 * vid[i] + ax] = rng.uniform(P.force_lo[ax], P.force_hi[ax]);
        }
      }
      // ---- AssignRandomMemberType
      for (int m = 0; m < M; ++m) {
        const int t = P.n_type > 1 ? rng.below(P.n_type) : 0;
        aed[3 * m] = a.type_table[3 * t];
        aed[3 * m + 1] = a.type_table[3 * t + 1];
        aed[3 * m + 2] = a.type_table[3 * t + 2];
      }
      // ---- the counting rule of Truss.isStable (truss.py:158-164): re-draw until it holds
      if (M + 3 * npin >= 3 * nJ) { status = 0; break; }
    }
    a.n_joint[o] = nJ;
    a.n_member[o] = M;
    a.info[o] = status;
  }
}

struct PackArgs {
  int n, max_joint, max_member;
  const double* xyz; const uint8_t* support; const double* force; const int32_t* conn; const double* aed;
  const int64_t* joint_off; const int64_t* member_off;
  double* oxyz; uint8_t* osup; double* oforce; int32_t* oconn; double* oaed;
};

__global__ void __launch_bounds__(128) k_gencube_pack(const PackArgs a) {
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int o = w; o < a.n; o += nw) {
    const int64_t j0 = a.joint_off[o], m0 = a.member_off[o];
    const int nJ = (int)(a.joint_off[o + 1] - j0), M = (int)(a.member_off[o + 1] - m0);
    const int64_t sj = (int64_t)o * a.max_joint, sm = (int64_t)o * a.max_member;
    for (int i = lane; i < 3 * nJ; i += 32) {
      a.oxyz[3 * j0 + i] = a.xyz[3 * sj + i];
      a.oforce[3 * j0 + i] = a.force[3 * sj + i];
    }
    for (int i = lane; i < nJ; i += 32) a.osup[j0 + i] = a.support[sj + i];
    for (int i = lane; i < 2 * M; i += 32) a.oconn[2 * m0 + i] = a.conn[2 * sm + i];
    for (int i = lane; i < 3 * M; i += 32) a.oaed[3 * m0 + i] = a.aed[3 * sm + i];
  }
}

}  // namespace

extern "C" int tb_gencube_limits(const tb_gencube_params* p, int32_t* max_joint, int32_t* max_member, int32_t* max_cube) {
  if (!p) return TB_ERR_NULL;
  const int gx = p->grid[0], gy = p->grid[1], gz = p->grid[2];
  if (gx < 1 || gy < 1 || gz < 1 || (int64_t)gx * gy * gz > TB_GEN_MAX_CELLS || (int64_t)(gx + 1) * (gy + 1) * (gz + 1) > TB_GEN_MAX_VERTS)
    return TB_ERR_TOO_LARGE;
  if (p->ncube_lo < 1 || p->ncube_hi < p->ncube_lo) return TB_ERR_SIZE;
  const int cells = gx * gy * gz, verts = (gx + 1) * (gy + 1) * (gz + 1);
  const int nc = p->ncube_hi < cells ? p->ncube_hi : cells;
  if (max_cube) *max_cube = nc;
  if (max_joint) *max_joint = 8 * nc < verts ? 8 * nc : verts;
  if (max_member) *max_member = 24 * nc;
  return TB_OK;
}

extern "C" int tb_gencube(const tb_gencube_params* p, int32_t n, const double* type_table, double* xyz, uint8_t* support,
                          double* force, int32_t* conn, double* aed, int32_t* n_joint, int32_t* n_member, int32_t* info,
                          int16_t* cells, uint8_t* picks, double* length, void* cuda_stream) {
  int32_t mj = 0, mm = 0, mc = 0;
  int rc = tb_gencube_limits(p, &mj, &mm, &mc);
  if (rc) return rc;
  if (n < 0) return TB_ERR_SIZE;
  if (p->method < 0 || p->method > 2 || p->link_type < 0 || p->link_type > 3 || p->n_type < 1 || p->max_attempts < 1) return TB_ERR_SIZE;
  if (!(p->length_hi >= p->length_lo)) return TB_ERR_SIZE;
  if (n == 0) return TB_OK;
  if (!type_table || !xyz || !support || !force || !conn || !aed || !n_joint || !n_member || !info) return TB_ERR_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return TB_ERR_NO_DEVICE;
  }
  GenArgs a;
  a.p = *p;
  a.n = n; a.max_joint = mj; a.max_member = mm; a.max_cube = mc;
  a.type_table = type_table;
  a.xyz = xyz; a.support = support; a.force = force; a.conn = conn; a.aed = aed;
  a.n_joint = n_joint; a.n_member = n_member; a.info = info;
  a.cells = cells; a.picks = picks; a.length = length;
  const int grid = (n + 63) / 64;
  k_gencube<<<grid, 64, 0, (cudaStream_t)cuda_stream>>>(a);
  tb_count_launch();
  return (int)cudaGetLastError();
}

extern "C" int tb_gencube_pack(const tb_gencube_params* p, int32_t n, const double* xyz, const uint8_t* support,
                               const double* force, const int32_t* conn, const double* aed, const int64_t* joint_off,
                               const int64_t* member_off, double* out_xyz, uint8_t* out_support, double* out_force,
                               int32_t* out_conn, double* out_aed, void* cuda_stream) {
  int32_t mj = 0, mm = 0;
  int rc = tb_gencube_limits(p, &mj, &mm, nullptr);
  if (rc) return rc;
  if (n < 0) return TB_ERR_SIZE;
  if (n == 0) return TB_OK;
  if (!xyz || !support || !force || !conn || !aed || !joint_off || !member_off || !out_xyz || !out_support || !out_force ||
      !out_conn || !out_aed)
    return TB_ERR_NULL;
  PackArgs a;
  a.n = n; a.max_joint = mj; a.max_member = mm;
  a.xyz = xyz; a.support = support; a.force = force; a.conn = conn; a.aed = aed;
  a.joint_off = joint_off; a.member_off = member_off;
  a.oxyz = out_xyz; a.osup = out_support; a.oforce = out_force; a.oconn = out_conn; a.oaed = out_aed;
  int grid = (n + 3) / 4;
  if (grid > 148 * 16) grid = 148 * 16;
  k_gencube_pack<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(a);
  tb_count_launch();
  return (int)cudaGetLastError();
}
