// Index gathers and their scatter-add gradients for sm_100a.
//
// Replaces gather_points / gather_points_grad (pointnet2/_ext_src/src/sampling.cpp:20-69, kernels
// sampling_gpu.cu:13-62) and group_points / group_points_grad (group_points.cpp:17-65, kernels
// group_points_gpu.cu:13-80).  The reference runs ONE block per scene (grid = B, or B x C for
// gather) so at B = 8 it leaves 140 of 148 SMs idle; here the (scene, channel-chunk, output-tile)
// space is flattened over the whole grid, each thread loads its index once and reuses it for a
// chunk of channels, and stores are unit-stride along the output's fastest axis.
//
// Forward results are pure copies (bit-exact).  The gradients are fp32 sums whose order the
// reference leaves to atomicAdd scheduling; here too (red.global.add.f32), so they are compared
// with a tolerance, as SURVEY.md 8(c) states.
#include "common.cuh"

namespace eda {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 8;  // channels per thread: one idx load feeds kChunk gathers

// points (B,C,N), idx (B,T) -> out (B,C,T).  T = M (gather) or M*S (group): group_points is the
// same gather with a two-level output index, group_points_gpu.cu:24-31 vs sampling_gpu.cu:17-23.
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N, int T, int cchunks,
                   float *__restrict__ out) {
  const int b = blockIdx.y / cchunks;
  const int c0 = (blockIdx.y % cchunks) * kChunk;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= T) return;
  const int i = __ldg(idx + (size_t)b * T + t);
  const float *__restrict__ src = points + ((size_t)b * C + c0) * N + i;
  float *__restrict__ dst = out + ((size_t)b * C + c0) * T + t;
  const int nc = min(kChunk, C - c0);
  float v[kChunk];
#pragma unroll
  for (int c = 0; c < kChunk; ++c)
    if (c < nc) v[c] = __ldg(src + (size_t)c * N);
#pragma unroll
  for (int c = 0; c < kChunk; ++c)
    if (c < nc) dst[(size_t)c * T] = v[c];
}

// grad_out (B,C,T), idx (B,T) -> grad_points (B,C,N) += ...   (pre-zeroed by the entry point)
__global__ void __launch_bounds__(kThreads)
scatter_rows_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C, int N, int T, int cchunks,
                    float *__restrict__ grad_points) {
  const int b = blockIdx.y / cchunks;
  const int c0 = (blockIdx.y % cchunks) * kChunk;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= T) return;
  const int i = __ldg(idx + (size_t)b * T + t);
  const float *__restrict__ src = grad_out + ((size_t)b * C + c0) * T + t;
  float *__restrict__ dst = grad_points + ((size_t)b * C + c0) * N + i;
  const int nc = min(kChunk, C - c0);
  float v[kChunk];
#pragma unroll
  for (int c = 0; c < kChunk; ++c)
    if (c < nc) v[c] = __ldg(src + (size_t)c * T);
#pragma unroll
  for (int c = 0; c < kChunk; ++c)
    if (c < nc) atomicAdd(dst + (size_t)c * N, v[c]);  // result unused -> RED.E.ADD.F32
}

int launch_gather(const float *points, const int *idx, int B, int C, int N, long long T, float *out, cudaStream_t st) {
  if (B < 0 || C < 0 || N < 0 || T < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || C == 0 || T == 0) return EDA_OK;
  if (!points || !idx || !out || N == 0) return EDA_ERR_INVALID_ARGUMENT;
  const int cchunks = (C + kChunk - 1) / kChunk;
  if (T > 0x7fffffffLL || (long long)B * cchunks > 65535) return EDA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((T + kThreads - 1) / kThreads), (unsigned)(B * cchunks));
  gather_rows_kernel<<<grid, kThreads, 0, st>>>(points, idx, C, N, (int)T, cchunks, out);
  return check_launch("gather_rows_kernel");
}

int launch_scatter(const float *grad_out, const int *idx, int B, int C, int N, long long T, float *grad_points,
                   cudaStream_t st) {
  if (B < 0 || C < 0 || N < 0 || T < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || C == 0 || N == 0) return EDA_OK;
  if (!grad_points) return EDA_ERR_INVALID_ARGUMENT;
  // sampling.cpp:52-54 / group_points.cpp:52-54: the reference returns a fresh zeros tensor
  EDA_CUDA_TRY(cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), st), "scatter memset");
  if (T == 0) return EDA_OK;
  if (!grad_out || !idx) return EDA_ERR_INVALID_ARGUMENT;
  const int cchunks = (C + kChunk - 1) / kChunk;
  if (T > 0x7fffffffLL || (long long)B * cchunks > 65535) return EDA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((T + kThreads - 1) / kThreads), (unsigned)(B * cchunks));
  scatter_rows_kernel<<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, (int)T, cchunks, grad_points);
  return check_launch("scatter_rows_kernel");
}

}  // namespace
}  // namespace eda

extern "C" {

int eda_gather_points(const float *points, const int *idx, int B, int C, int N, int M, float *out, void *stream) {
  return eda::launch_gather(points, idx, B, C, N, M, out, eda::as_stream(stream));
}

int eda_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, float *grad_points,
                           void *stream) {
  return eda::launch_scatter(grad_out, idx, B, C, N, M, grad_points, eda::as_stream(stream));
}

int eda_group_points(const float *points, const int *idx, int B, int C, int N, int M, int S, float *out,
                     void *stream) {
  if (M < 0 || S < 0) return EDA_ERR_INVALID_ARGUMENT;
  return eda::launch_gather(points, idx, B, C, N, (long long)M * S, out, eda::as_stream(stream));
}

int eda_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, int S,
                          float *grad_points, void *stream) {
  if (M < 0 || S < 0) return EDA_ERR_INVALID_ARGUMENT;
  return eda::launch_scatter(grad_out, idx, B, C, N, (long long)M * S, grad_points, eda::as_stream(stream));
}

}  // extern "C"
