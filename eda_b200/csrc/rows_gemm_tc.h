// Internal interface of the persistent tcgen05 row-streaming GEMM (rows_gemm_tc.cu), used by eda_rows_gemm_stats.
#pragma once
#include <cuda_runtime.h>

namespace eda {
bool rows_gemm_tc_eligible(const float *x, int ldx, long long rows, int K, int N, const float *y, int ldy);
constexpr int kRowsGemmTcDeclined = -1000;  // nothing launched (tensor map could not be encoded / does not fit)
int rows_gemm_tc_launch(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                        long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y, int ldy,
                        double *stats, cudaStream_t stream);
}  // namespace eda
