// Launch logic of the trajectory-covariance kernels (fg_cov.cuh).
#include <cstdlib>

#include "fg_cov.cuh"
#include "fg_cov_umma.cuh"
#include "fg_dispatch.h"

namespace fg {

template <typename K>
static cudaError_t opt_in(K kern, size_t smem) {
  if (smem > 48 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

// 3-D tensor map of the output: (column, time point t1, slab = env * dof + d1), box 32 x 32 x 1 floats, 128B swizzle.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
static cudaError_t make_out_map(CUtensorMap* map, float* cov, int NT, int T, long long slabs, const char** why) {
  typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_t encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || !fn) {
      *why = "cuTensorMapEncodeTiled is not available in this driver";
      return cudaSuccess;
    }
    encode = (encode_t)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)NT, (cuuint64_t)T, (cuuint64_t)slabs};
  const cuuint64_t strides[2] = {(cuuint64_t)NT * 4, (cuuint64_t)T * NT * 4};
  const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, cov, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) *why = "cuTensorMapEncodeTiled rejected the output layout (needs 16-byte aligned rows)";
  return cudaSuccess;
}

cudaError_t launch_traj_cov(const CovArgs& a0, long long B, int path, cudaStream_t stream, int max_smem_optin,
                            const char** why) {
  CovArgs a = a0;
  const int D = a.N * a.Kc, NT = a.N * a.T;
  if (D > kCovMaxD || a.Kc > 16) {
    *why = "dof * basis count too large for the covariance kernels (<= 96, basis <= 16)";
    return cudaSuccess;
  }
  cudaError_t e = cudaMemsetAsync(a.gmax, 0, sizeof(float), stream);
  if (e != cudaSuccess) return e;
  const size_t smem_diag = sizeof(float) * ((size_t)D * D + (size_t)a.T * a.Kc);
  if (smem_diag > (size_t)max_smem_optin) {
    *why = "covariance tables exceed the shared memory of one SM";
    return cudaSuccess;
  }
  if ((e = opt_in(k_cov_diag, smem_diag)) != cudaSuccess) return e;
  k_cov_diag<<<(unsigned)B, kCovThreads, smem_diag, stream>>>(a);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (a.cov) {
    if (path == 2) {
      const UmmaSmem sp = umma_smem_plan(a.Kc, a.N, a.T);
      if (sp.total > max_smem_optin || (NT & 3) != 0 || B > 65535) {
        *why = "tcgen05 covariance path: operands exceed shared memory, rows are not 16-byte aligned, or batch > 65535";
        return cudaSuccess;
      }
      CUtensorMap map;
      if ((e = make_out_map(&map, a.cov, NT, a.T, B * a.N, why)) != cudaSuccess || *why) return e;
      if ((e = opt_in(k_cov_umma, (size_t)sp.total)) != cudaSuccess) return e;
      const dim3 grid((a.T + kUmmaM - 1) / kUmmaM, a.N, (unsigned)B);
      k_cov_umma<<<grid, kUmmaThreads, sp.total, stream>>>(a, map);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    } else {
    auto pad4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    const size_t smem = sizeof(float) * (pad4((size_t)D * D) + pad4((size_t)a.T * a.Kc) + pad4((size_t)a.Kc * D) + pad4((size_t)a.Kc * NT) +
                                         2 * (size_t)a.T * ((a.Kc + 1) & ~1));
    if (smem > (size_t)max_smem_optin) {
      *why = "covariance slab exceeds the shared memory of one SM";
      return cudaSuccess;
    }
    // rows per block: ~192 KB of output per block.  Measured (profiles/README.md): smaller blocks pay the slab set-up too
    // often, larger ones leave the hardware block scheduler too little to balance the store stream with (1000 x 1000:
    // 50 rows 6.05 TB/s, 100 rows 4.8, 200 rows 4.4; 400 x 400: 50 rows 4.8, 100 rows 5.9, 200 rows 4.3).
    {
      const int target = (int)((192 * 1024) / ((size_t)NT * 4));
      const int chunks = (a.T + (target > 8 ? target : 8) - 1) / (target > 8 ? target : 8);
      a.rows_per_block = (a.T + chunks - 1) / chunks;
    }
    if (const char* ev = getenv("FG_COV_ROWS")) if (atoi(ev) > 0) a.rows_per_block = atoi(ev);
    const dim3 grid((a.T + a.rows_per_block - 1) / a.rows_per_block, a.N, (unsigned)B);
    if (B > 65535) {
      *why = "covariance batch larger than 65535 envs per call";
      return cudaSuccess;
    }
#define FG_COV(KC)                                                  \
  {                                                                 \
    if ((e = opt_in(k_cov_simt<KC>, smem)) != cudaSuccess) return e; \
    k_cov_simt<KC><<<grid, kCovThreads, smem, stream>>>(a);         \
  }
    if (a.Kc == 5) FG_COV(5) else if (a.Kc == 6) FG_COV(6) else FG_COV(0)
#undef FG_COV
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
  }
  if (a.stdv) {
    const long long n = B * a.T * a.N;
    k_cov_std<<<(unsigned)((n + kCovThreads - 1) / kCovThreads), kCovThreads, 0, stream>>>(a, B);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace fg
