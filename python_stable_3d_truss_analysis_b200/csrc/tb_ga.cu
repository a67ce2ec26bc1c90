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

// One CTA ranks the population: bitonic sort of (fitness key, index) pairs in shared memory -- the index as the second
// key makes it the stable sort of `sorted(pop, key=fitness)` (ga.py:157).  Then the report: best individual and the
// first feasible one in rank order (_RecordFeasible, ga.py:101-108).
__global__ void __launch_bounds__(1024) k_ga_rank(int n_pop, int p2, const double* fitness, const uint8_t* flags, int32_t* order,
                                                  tb_ga_report* rep) {
  extern __shared__ __align__(16) unsigned char smraw[];
  uint64_t* key = reinterpret_cast<uint64_t*>(smraw);
  int32_t* idx = reinterpret_cast<int32_t*>(key + p2);
  __shared__ int first_feasible;
  const int tid = threadIdx.x;
  for (int i = tid; i < p2; i += 1024) {
    key[i] = i < n_pop ? key_of(fitness[i]) : ~0ull;
    idx[i] = i < n_pop ? i : 0x7fffffff;
  }
  if (tid == 0) first_feasible = 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < p2; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const uint64_t ka = key[i], kb = key[l];
          const int ia = idx[i], ib = idx[l];
          const bool gt = ka > kb || (ka == kb && ia > ib);
          if (gt == up) {
            key[i] = kb; key[l] = ka;
            idx[i] = ib; idx[l] = ia;
          }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < n_pop; i += 1024) {
    const int g = idx[i];
    order[i] = g;
    if (flags && flags[2 * g] && flags[2 * g + 1]) atomicMin(&first_feasible, i);
  }
  __syncthreads();
  if (tid == 0 && rep) {
    const int b = idx[0];
    rep->best_index = b;
    rep->best_fitness = fitness[b];
    rep->best_stress_ok = flags ? flags[2 * b] : 0;
    rep->best_displace_ok = flags ? flags[2 * b + 1] : 0;
    rep->feasible_index = first_feasible == 0x7fffffff ? -1 : idx[first_feasible];
    rep->feasible_fitness = first_feasible == 0x7fffffff ? 0.0 : fitness[idx[first_feasible]];
  }
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
  int p2 = 1;
  while (p2 < p->n_pop) p2 <<= 1;
  const size_t smem = (size_t)p2 * 12;
  if (smem > 200 * 1024) return TB_ERR_TOO_LARGE;                              // one-CTA ranking: up to 16384 individuals
  static size_t granted = 0;
  if (granted < smem) {
    cudaError_t e = cudaFuncSetAttribute(k_ga_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    granted = smem;
  }
  k_ga_rank<<<1, 1024, smem, st>>>(p->n_pop, p2, fitness, flags, order, report);
  int launches = 1;
  if (gene_out) {                              // gene_out == NULL: ranking and report only (the final Select)
    const int64_t total = (int64_t)p->n_pop * p->n_member;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_ga_update<<<grid, 256, 0, st>>>(*p, (uint32_t)generation, gene_in, order, gene_out);
    ++launches;
  }
  tb_count_launch(launches);
  return (int)cudaGetLastError();
}
