// Dataset generation on the device: a pool of trusses (packed ragged arrays, e.g. the output
// of GenerateRandomCubeTrusses) is expanded into n_out augmented trusses, ready for tb_solve_ragged -- the augmenters of
// slientruss3d/generate.py applied to packed arrays, one warp per output truss:
//   MoveToCentroid     :62-76   centroid = (joint positions summed in joint order) / nJ, subtracted (same roundings)
//   RandomTranslation  :97-106  one uniform vector per truss, added to every joint
//   AddJointNoise      :43-59   independent Gaussian noise per joint coordinate
//   RandomResetPin     :109-132 k uniform in [max(minNumPin, ceil((3 nJ - M) / 3)), int(ratio nJ)], k random joints PIN, the rest NO
// applied in this order (the order of the reference's recipe, example.py:239-267; each may be switched off).
// Random numbers: counter-based Philox4x32-10 keyed by the seed, counter = (output truss, element, purpose): reproducible
// and order-free.  The host augmenter classes keep Python's `random` stream; this is the bulk path.
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
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

enum : uint32_t { PUR_TRANSLATE = 11, PUR_NOISE = 12, PUR_PINCOUNT = 13, PUR_PINKEY = 14 };

struct AugArgs {
  int dim, n_out;
  const int64_t* pj; const int64_t* pm;       // pool offsets
  const double* pxyz; const uint8_t* psup; const int32_t* pconn; const double* paed; const double* pforce;
  const int32_t* src;
  const int64_t* oj; const int64_t* om;       // output offsets
  double* oxyz; uint8_t* osup; int32_t* oconn; double* oaed; double* oforce;
  tb_augment_params prm;
  int max_joint;
};

__global__ void __launch_bounds__(128) k_augment(const AugArgs a) {
  extern __shared__ uint32_t sKeys[];          // [4 warps][max_joint] pin-selection keys
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* keys = sKeys + (size_t)wid * a.max_joint;
  const int d = a.dim;
  for (int o = blockIdx.x * 4 + wid; o < a.n_out; o += gridDim.x * 4) {
    const int s = a.src[o];
    const int64_t j0 = a.pj[s], m0 = a.pm[s];
    const int nJ = (int)(a.pj[s + 1] - j0), M = (int)(a.pm[s + 1] - m0);
    const int64_t J0 = a.oj[o], M0 = a.om[o];
    // members, loads
    for (int i = lane; i < 2 * M; i += 32) a.oconn[2 * M0 + i] = a.pconn[2 * m0 + i];
    for (int i = lane; i < 3 * M; i += 32) a.oaed[3 * M0 + i] = a.paed[3 * m0 + i];
    for (int i = lane; i < d * nJ; i += 32) a.oforce[d * J0 + i] = a.pforce[d * j0 + i];
    // joint positions
    double cen = 0.0;
    if (a.prm.move_to_centroid && lane < d) {  // GetCentroid, generate.py:23-28: sequential sum in joint order, / n
      double sum = 0.0;
      for (int j = 0; j < nJ; ++j) sum = __dadd_rn(sum, a.pxyz[d * (j0 + j) + lane]);
      cen = __ddiv_rn(sum, (double)nJ);
    }
    double c3[3], t3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c3[i] = __shfl_sync(0xffffffffu, cen, i);
      t3[i] = 0.0;
    }
    if (a.prm.random_translation) {
      const U4 r0 = philox((uint32_t)o, 0u, 0u, PUR_TRANSLATE, a.prm.seed), r1 = philox((uint32_t)o, 1u, 0u, PUR_TRANSLATE, a.prm.seed);
      const double w = a.prm.translate_hi - a.prm.translate_lo;                  // random.uniform(a, b) = a + (b - a) * random()
      t3[0] = a.prm.translate_lo + w * u01(r0.x, r0.y);
      t3[1] = a.prm.translate_lo + w * u01(r0.z, r0.w);
      t3[2] = a.prm.translate_lo + w * u01(r1.x, r1.y);
    }
    for (int e = lane; e < d * nJ; e += 32) {
      const int ax = e % d;
      double x = a.pxyz[d * j0 + e];
      if (a.prm.move_to_centroid) x = __dsub_rn(x, c3[ax]);
      if (a.prm.random_translation) x = __dadd_rn(x, t3[ax]);
      if (a.prm.joint_noise) {
        const U4 r = philox((uint32_t)o, (uint32_t)e, 0u, PUR_NOISE, a.prm.seed);
        const double u1 = 1.0 - u01(r.x, r.y), u2 = u01(r.z, r.w);               // u1 in (0, 1]
        const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);                 // Box-Muller
        x = __dadd_rn(x, a.prm.noise_mean[ax] + a.prm.noise_std[ax] * z);
      }
      a.oxyz[d * J0 + e] = x;
    }
    // supports
    if (!a.prm.reset_pin) {
      for (int j = lane; j < nJ; j += 32) a.osup[J0 + j] = a.psup[j0 + j];
    } else {
      int need = (3 * nJ - M + 2) / 3;                                           // ceil((3 nJ - M) / 3), GetStableMinNumPin
      if (3 * nJ - M < 0) need = -((M - 3 * nJ) / 3);
      int lo = a.prm.min_pin > need ? a.prm.min_pin : need;
      int hi = a.prm.max_pin_ratio > 0.0 ? (int)(a.prm.max_pin_ratio * nJ) : nJ;
      if (hi > nJ) hi = nJ;
      if (lo > nJ) lo = nJ;
      if (hi < lo) hi = lo;                                                      // the reference raises here (empty range)
      const U4 rc = philox((uint32_t)o, 0u, 0u, PUR_PINCOUNT, a.prm.seed);
      const int k = lo + (int)(((uint64_t)rc.x * (uint64_t)(hi - lo + 1)) >> 32);
      __syncwarp();
      for (int j = lane; j < nJ; j += 32) keys[j] = philox((uint32_t)o, (uint32_t)j, 0u, PUR_PINKEY, a.prm.seed).x;
      __syncwarp();
      for (int j = lane; j < nJ; j += 32) {                                      // the k smallest (key, joint) pairs are the pins
        const uint32_t kj = keys[j];
        int rank = 0;
        for (int q = 0; q < nJ; ++q) rank += (keys[q] < kj) || (keys[q] == kj && q < j);
        a.osup[J0 + j] = rank < k ? SUP_PIN : SUP_NO;
      }
      __syncwarp();
    }
  }
}

}  // namespace

extern "C" int tb_augment_ragged(const tb_ragged_in* pool, int32_t n_out, const int32_t* src, const int64_t* out_joint_off,
                                 const int64_t* out_member_off, const tb_augment_params* prm, double* out_xyz,
                                 uint8_t* out_support, int32_t* out_conn, double* out_aed, double* out_force, void* cuda_stream) {
  if (!pool || !prm) return TB_ERR_NULL;
  if (pool->dim != 2 && pool->dim != 3) return TB_ERR_DIM;
  if (n_out < 0 || pool->batch <= 0 || pool->max_joint <= 0) return TB_ERR_SIZE;
  if (n_out == 0) return TB_OK;
  if (!pool->joint_off || !pool->member_off || !pool->joint_xyz || !pool->support || !pool->conn || !pool->member_aed ||
      !pool->force || !src || !out_joint_off || !out_member_off || !out_xyz || !out_support || !out_conn || !out_aed || !out_force)
    return TB_ERR_NULL;
  if (prm->reset_pin && prm->min_pin < 3) return TB_ERR_SIZE;                    // PinNotEnoughError, generate.py:113-114
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return TB_ERR_NO_DEVICE;
  AugArgs a;
  a.dim = pool->dim;
  a.n_out = n_out;
  a.pj = pool->joint_off; a.pm = pool->member_off;
  a.pxyz = pool->joint_xyz; a.psup = pool->support; a.pconn = pool->conn; a.paed = pool->member_aed; a.pforce = pool->force;
  a.src = src;
  a.oj = out_joint_off; a.om = out_member_off;
  a.oxyz = out_xyz; a.osup = out_support; a.oconn = out_conn; a.oaed = out_aed; a.oforce = out_force;
  a.prm = *prm;
  a.max_joint = pool->max_joint;
  const size_t smem = (size_t)4 * pool->max_joint * sizeof(uint32_t);
  if (smem > 96 * 1024) return TB_ERR_TOO_LARGE;
  static size_t granted = 48 * 1024;
  if (smem > granted) {
    cudaError_t e = cudaFuncSetAttribute(k_augment, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    granted = smem;
  }
  int grid = (n_out + 3) / 4;
  if (grid > 148 * 16) grid = 148 * 16;
  k_augment<<<grid, 128, smem, (cudaStream_t)cuda_stream>>>(a);
  tb_count_launch(1);
  return (int)cudaGetLastError();
}
