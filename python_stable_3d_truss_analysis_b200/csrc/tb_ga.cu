// GA generation step on the device: ranking, elitism, crossover, mutation and re-seeding of a
// whole population without leaving the GPU, so that a generation is tb_fitness + tb_ga_step and the host reads back
// 40 bytes.  The operators are the reference's (slientruss3d/ga.py):
//   Select     :155-160  sort by fitness (stable: ties keep population order), the first nElite genes are the elites
//   UpdatePop  :172-190  child j >= nElite draws p: <= pCrossover -> Crossover of two distinct elites,
//                        <= +pMutate -> Mutate of one elite, <= +pOrigin -> keeps pop[j], else a fresh random gene
//   Crossover  :162-165  two distinct cut points, the child takes gene1 on [cut0, cut1) and gene0 elsewhere
//   Mutate     :167-171  one position, replaced by a different member type
//   Initialize :151-153  every gene drawn from the (weighted) member-type distribution
// Random numbers: counter-based Philox4x32-10 keyed by the seed, counter = (individual, generation, purpose) -- any
// thread can reproduce any decision, so one thread per (individual, member) needs no communication.  The reference
// draws from Python's Mersenne Twister in program order, which cannot be replayed in parallel; the host GA
// (python_stable_3d_truss_analysis_b200/ga.py: GA.Evolve) stays stream-compatible with it, this path is the fast one.
#include <mutex>

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
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {   // 53 random bits -> [0, 1)
  return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ int below(uint32_t r, int n) { return (int)(((uint64_t)r * (uint64_t)n) >> 32); }   // bias < n / 2^32

enum : uint32_t { PUR_BRANCH = 1, PUR_GENE = 2, PUR_INIT = 3 };

__device__ __forceinline__ int draw_type(uint32_t a, uint32_t b, const double* cum, int n_type) {
  if (!cum) return below(a, n_type);
  const double u = u01(a, b);                  // cum[t] = normalised cumulative weight, cum[n_type-1] = 1
  int t = 0;
  while (t < n_type - 1 && u >= cum[t]) ++t;
  return t;
}

__global__ void k_ga_init(tb_ga_params p, const double* cum, int32_t* gene) {
  const int64_t total = (int64_t)p.n_pop * p.n_member;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const U4 r = philox((uint32_t)i, (uint32_t)(i >> 32), 0u, PUR_INIT, p.seed);
    gene[i] = draw_type(r.x, r.y, cum, p.n_type);
  }
}

// order-preserving map of a double onto uint64 (total order; -0 < +0; NaN sorts last among positives)
__device__ __forceinline__ uint64_t key_of(double f) {
  const uint64_t b = (uint64_t)__double_as_longlong(f);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Ranking without a sort: the rank of individual i is the number of individuals that precede it in the stable order of
// `sorted(pop, key=fitness)` (ga.py:157), i.e. #{j : key_j < key_i or (key_j == key_i and j < i)}.  N^2 comparisons are
// nothing for a GPU (8192^2 = 67 M), they spread over a 2-D grid (tiles of 256 individuals x slices of the comparison
// range, partial counts added atomically), and there is no one-CTA bottleneck and no population limit.
__global__ void __launch_bounds__(256) k_ga_rank_count(int n_pop, int slice, const double* __restrict__ fitness, int32_t* __restrict__ rank) {
  __shared__ uint64_t sKey[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const uint64_t ki = i < n_pop ? key_of(fitness[i]) : 0ull;
  const int j0 = blockIdx.y * slice, j1 = min(n_pop, j0 + slice);
  int cnt = 0;
  for (int base = j0; base < j1; base += 256) {
    const int j = base + threadIdx.x;
    __syncthreads();
    sKey[threadIdx.x] = j < j1 ? key_of(fitness[j]) : ~0ull;
    __syncthreads();
    const int m = min(256, j1 - base);
#pragma unroll 8
    for (int t = 0; t < m; ++t) {
      const uint64_t kj = sKey[t];
      cnt += (kj < ki) || (kj == ki && base + t < i);
    }
  }
  if (i < n_pop && cnt) atomicAdd(&rank[i], cnt);
}

// order[rank] = individual; the first feasible individual in rank order = the smallest rank among the feasible ones
__global__ void __launch_bounds__(256) k_ga_rank_scatter(int n_pop, const int32_t* __restrict__ rank, const uint8_t* __restrict__ flags,
                                                          int32_t* __restrict__ order, int32_t* __restrict__ first_feasible) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_pop) return;
  const int r = rank[i];
  order[r] = i;
  if (flags && flags[2 * i] && flags[2 * i + 1]) atomicMin(first_feasible, r);
}

// the report: best individual and the first feasible one in rank order (_RecordFeasible, ga.py:101-108)
__global__ void k_ga_report(const double* fitness, const uint8_t* flags, const int32_t* order, const int32_t* first_feasible, tb_ga_report* rep) {
  const int b = order[0], ff = *first_feasible;
  rep->best_index = b;
  rep->best_fitness = fitness[b];
  rep->best_stress_ok = flags ? flags[2 * b] : 0;
  rep->best_displace_ok = flags ? flags[2 * b + 1] : 0;
  rep->feasible_index = ff >= 0x7f7f7f7f ? -1 : order[ff];
  rep->feasible_fitness = ff >= 0x7f7f7f7f ? 0.0 : fitness[order[ff]];
}

// One thread per (individual, member).  The decisions of individual j come from Philox block (j, generation, BRANCH):
// x,y -> p; z -> first parent, w -> second parent / mutation position; a second block gives the cut points and the new type.
__global__ void k_ga_update(tb_ga_params p, uint32_t generation, const int32_t* gene_in, const int32_t* order, int32_t* gene_out) {
  const int M = p.n_member;
  const int64_t total = (int64_t)p.n_pop * M;
  const double e_cross = p.p_crossover, e_mut = e_cross + p.p_mutate, e_orig = e_mut + p.p_origin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i / M), m = (int)(i - (int64_t)j * M);
    int v;
    if (j < p.n_elite) {
      v = gene_in[(int64_t)order[j] * M + m];                       // newPop[:nElite] = elitePop, ga.py:178
    } else {
      const U4 r = philox((uint32_t)j, generation, 0u, PUR_BRANCH, p.seed);
      const double pr = u01(r.x, r.y);
      if (pr <= e_cross) {
        const U4 s = philox((uint32_t)j, generation, 1u, PUR_BRANCH, p.seed);
        int a = below(r.z, p.n_elite), b = below(r.w, p.n_elite - 1);   // random.sample(elitePop, k=2): distinct, ordered
        if (b >= a) ++b;
        int c0 = below(s.x, M), c1 = below(s.y, M - 1);                 // random.sample(range(nMember), k=2)
        if (c1 >= c0) ++c1;
        const int lo = min(c0, c1), hi = max(c0, c1);
        const int src = (m < lo || m >= hi) ? a : b;
        v = gene_in[(int64_t)order[src] * M + m];
      } else if (pr <= e_mut) {
        const U4 s = philox((uint32_t)j, generation, 1u, PUR_BRANCH, p.seed);
        const int a = below(r.z, p.n_elite);                            // random.choice(elitePop)
        const int at = below(r.w, M);                                   // random.randint(0, nMember - 1)
        v = gene_in[(int64_t)order[a] * M + m];
        if (m == at) {
          int t = below(s.x, p.n_type - 1);                             // a different member type
          if (t >= v) ++t;
          v = t;
        }
      } else if (pr <= e_orig) {
        v = gene_in[i];                                                 // newPop[j] = pop[j]
      } else {
        const U4 s = philox((uint32_t)i, (uint32_t)(i >> 32) ^ (generation << 8), 2u, PUR_GENE, p.seed);
        v = below(s.x, p.n_type);                                       // GetRandomGene: uniform, ga.py:129-130
      }
    }
    gene_out[i] = v;
  }
}

int check_params(const tb_ga_params* p) {
  if (!p) return TB_ERR_NULL;
  if (p->n_pop <= 0 || p->n_member <= 0 || p->n_type < 2 || p->n_elite < 0 || p->n_elite > p->n_pop) return TB_ERR_SIZE;
  if (p->p_crossover < 0 || p->p_mutate < 0 || p->p_origin < 0 || p->p_crossover + p->p_mutate + p->p_origin > 1.0 + 1e-12)
    return TB_ERR_SIZE;
  return TB_OK;
}

}  // namespace

extern "C" int tb_ga_init(const tb_ga_params* p, const double* type_cum, int32_t* gene, void* cuda_stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!gene) return TB_ERR_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  const int64_t total = (int64_t)p->n_pop * p->n_member;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  k_ga_init<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(*p, type_cum, gene);
  tb_count_launch(1);
  return (int)cudaGetLastError();
}

extern "C" int tb_ga_step(const tb_ga_params* p, uint64_t generation, const double* fitness, const uint8_t* flags,
                          const int32_t* gene_in, int32_t* gene_out, int32_t* order, tb_ga_report* report, void* cuda_stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!fitness || !gene_in || !order) return TB_ERR_NULL;
  if (gene_out && p->n_elite < 2 && p->p_crossover > 0) return TB_ERR_SIZE;   // crossover needs two distinct elites
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // scratch of the ranking: rank[n_pop] + the smallest feasible rank, one grow-only buffer per device
  static std::mutex scratch_mu;
  static int32_t* scratch[64] = {};
  static size_t scratch_cap[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  int32_t* rank = nullptr;
  {
    std::lock_guard<std::mutex> lock(scratch_mu);
    const size_t need = (size_t)p->n_pop + 1;
    if (scratch_cap[dev & 63] < need) {
      if (scratch[dev & 63]) {
        cudaDeviceSynchronize();
        cudaFree(scratch[dev & 63]);
        scratch[dev & 63] = nullptr;
        scratch_cap[dev & 63] = 0;
      }
      cudaError_t e = cudaMalloc((void**)&scratch[dev & 63], need * sizeof(int32_t));
      if (e != cudaSuccess) return (int)e;
      scratch_cap[dev & 63] = need;
    }
    rank = scratch[dev & 63];
  }
  int32_t* first_feasible = rank + p->n_pop;
  {
    cudaError_t e = cudaMemsetAsync(rank, 0, (size_t)p->n_pop * sizeof(int32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(first_feasible, 0x7f, sizeof(int32_t), st);   // 0x7f7f7f7f: "none" (> any rank)
    if (e != cudaSuccess) return (int)e;
  }
  const int tiles = (p->n_pop + 255) / 256;
  int splits = (148 * 8 + tiles - 1) / tiles;                                    // enough CTAs to fill the GPU
  if (splits > tiles) splits = tiles;
  if (splits < 1) splits = 1;
  const int slice = ((p->n_pop + splits - 1) / splits + 255) / 256 * 256;
  k_ga_rank_count<<<dim3(tiles, (p->n_pop + slice - 1) / slice), 256, 0, st>>>(p->n_pop, slice, fitness, rank);
  k_ga_rank_scatter<<<tiles, 256, 0, st>>>(p->n_pop, rank, flags, order, first_feasible);
  if (report) k_ga_report<<<1, 1, 0, st>>>(fitness, flags, order, first_feasible, report);
  int launches = report ? 3 : 2;
  if (gene_out) {                              // gene_out == NULL: ranking and report only (the final Select)
    const int64_t total = (int64_t)p->n_pop * p->n_member;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_ga_update<<<grid, 256, 0, st>>>(*p, (uint32_t)generation, gene_in, order, gene_out);
    ++launches;
  }
  tb_count_launch(launches);
  return (int)cudaGetLastError();
}
