// 2-D TMA tensor maps without a link-time dependency on libcuda: cuTensorMapEncodeTiled is fetched through the runtime's
// driver entry point table.  Shared by the forward GEMM (linear.cu) and the tcgen05 weight-gradient kernel (wgrad_tc.cu).
#pragma once
#include <stdlib.h>
#include <cuda.h>  // CUtensorMap (types only)
#include "common.cuh"

namespace eda {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// nullptr when the driver entry point is unavailable or EDA_LINEAR_TMA=0 (callers then use their non-TMA path)
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    if (const char *e = getenv("EDA_LINEAR_TMA")) if (e[0] == '0') return nullptr;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// Map over a row-major fp32 matrix (rows, cols; row stride ld floats) for boxes of box_rows x 32 columns: one box row =
// one 128-byte shared-memory row.  atom32 = false: SWIZZLE_128B, 16-byte chunk j of row r at chunk position j ^ (r & 7)
// (K-major tensor-core operands); atom32 = true: SWIZZLE_128B_ATOM_32B, 32-byte unit u of row r at unit position
// u ^ (r & 3) (what MN-major kind::tf32 operands need: UMMA layout type 1, eda_selftest_umma_probe).  Elements outside
// the matrix read as zero.  false = cannot be encoded (no driver entry point, misaligned base / stride).
inline bool make_tensor_map_rows32(CUtensorMap *map, const float *x, long long rows, int cols, long long ld, int box_rows,
                                   bool atom32 = false) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || (ld & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || rows <= 0 || cols <= 0 || box_rows < 1 || box_rows > 256)
    return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), gdim, gstride, box, estride,
             CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D tiled TMA load global -> shared (tensor map in kernel-parameter space), completes on `bar` with the box bytes
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace eda
