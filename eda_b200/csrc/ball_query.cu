// Ball query for sm_100a.
//
// Replaces ball_query (pointnet2/_ext_src/src/ball_query.cpp:13-37; kernel
// ball_query_gpu.cu:14-49): for every centre, the first `nsample` point indices in ascending
// index order with d2 < radius^2 (strict, f32 radius*radius), remaining slots padded with the
// first hit, all zeros for an empty ball.
//
// The reference gives each centre to ONE thread of ONE block per scene and walks xyz in
// global memory.  Here a CTA owns kWarps centres and streams the scene's xyz through shared
// memory in tiles fetched by the TMA engine (1-D cp.async.bulk, double buffered, mbarrier
// completion).  One warp per centre tests 128 points per step (4 per lane); hits are rare, so
// the common step is branch-free, and the ordered compaction (ballot + prefix popcount)
// keeps "first nsample in index order" exact.  The row is staged in smem and written once,
// coalesced, including the padding — idx needs no zero-fill by the caller.
#include "common.cuh"

namespace eda {
namespace {

constexpr int kBqWarps = 16;
constexpr int kBqThreads = kBqWarps * 32;
constexpr int kTile = 2048;  // points per smem tile: 24 KB, x2 buffers
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float *__restrict__ new_xyz_all, const float *__restrict__ xyz_all, int N, int M, float r2,
                  int nsample, int use_tma, int *__restrict__ idx_all) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile0 = reinterpret_cast<float *>(smem_raw);
  float *tile1 = tile0 + kTile * 3;
  int *hits = reinterpret_cast<int *>(tile1 + kTile * 3);  // [kBqWarps][nsample]
  __shared__ __align__(8) uint64_t full[2];

  const int b = blockIdx.y;
  const float *__restrict__ xyz = xyz_all + (size_t)b * N * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int centre = blockIdx.x * kBqWarps + warp;
  const bool active = centre < M;
  int *myhits = hits + warp * nsample;

  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (active) {
    const float *c = new_xyz_all + ((size_t)b * M + centre) * 3;
    cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
  }
  const int ntiles = (N + kTile - 1) / kTile;
  auto tile_pts = [&](int t) { return min(kTile, N - t * kTile); };

  if (use_tma) {
    if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      for (int t = 0; t < 2 && t < ntiles; ++t) {
        const uint32_t bytes = (uint32_t)tile_pts(t) * 12u;
        mbar_arrive_expect_tx(&full[t], bytes);
        bulk_g2s(t ? tile1 : tile0, xyz + (size_t)t * kTile * 3, bytes, &full[t]);
      }
    }
  }

  int cnt = 0;  // warp-uniform number of hits so far (may exceed nsample in the last step)
  bool done = !active;
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    float *tile = buf ? tile1 : tile0;
    const int npts = tile_pts(t);
    if (use_tma) {
      mbar_wait(&full[buf], (t >> 1) & 1);
    } else {
      // unaligned scenes (N % 4 != 0): plain coalesced loads into the same tile
      const float *src = xyz + (size_t)t * kTile * 3;
      for (int i = tid; i < npts * 3; i += kBqThreads) tile[i] = __ldg(src + i);
      __syncthreads();
    }
    if (!done) {
      const int kbase = t * kTile;
      for (int base = 0; base < npts && cnt < nsample; base += 128) {
        bool hit[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = base + u * 32 + lane;
          const int kk = min(k, npts - 1);
          const float x = tile[kk * 3 + 0], y = tile[kk * 3 + 1], z = tile[kk * 3 + 2];
          // (new_x - x)^2 + ... as the reference compiles it: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
          const float d2 = sq3(__fsub_rn(cx, x), __fsub_rn(cy, y), __fsub_rn(cz, z));
          hit[u] = (k < npts) && (d2 < r2);
        }
        if (__any_sync(kFull, hit[0] | hit[1] | hit[2] | hit[3])) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const unsigned mask = __ballot_sync(kFull, hit[u]);
            if (hit[u]) {
              const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
              if (pos < nsample) myhits[pos] = kbase + base + u * 32 + lane;
            }
            cnt += __popc(mask);
          }
        }
      }
      done = cnt >= nsample;
    }
    // every warp is past this tile -> its buffer can be refilled; stop early when all are done
    const int all_done = __syncthreads_and(done ? 1 : 0);
    if (all_done) {
      if (use_tma && t + 1 < ntiles) mbar_wait(&full[(t + 1) & 1], ((t + 1) >> 1) & 1);  // drain in-flight copy
      break;
    }
    if (use_tma && tid == 0 && t + 2 < ntiles) {
      const uint32_t bytes = (uint32_t)tile_pts(t + 2) * 12u;
      mbar_arrive_expect_tx(&full[buf], bytes);
      bulk_g2s(tile, xyz + (size_t)(t + 2) * kTile * 3, bytes, &full[buf]);
    }
  }

  if (active) {
    __syncwarp();
    int *out = idx_all + ((size_t)b * M + centre) * nsample;
    const int have = min(cnt, nsample);
    const int first = have > 0 ? myhits[0] : 0;  // ball_query_gpu.cu:38-42 pads with the first hit
    for (int s = lane; s < nsample; s += 32) out[s] = s < have ? myhits[s] : first;
  }
}

}  // namespace
}  // namespace eda

extern "C" int eda_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius, int nsample,
                              int *idx, void *stream) {
  using namespace eda;
  if (B < 0 || N < 0 || M < 0 || nsample < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0 || nsample == 0) return EDA_OK;
  if (!new_xyz || !idx || (N > 0 && !xyz)) return EDA_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {  // nothing to find: the reference returns its zero-initialised tensor
    EDA_CUDA_TRY(cudaMemsetAsync(idx, 0, (size_t)B * M * nsample * sizeof(int), st), "ball_query memset");
    return EDA_OK;
  }
  const size_t smem = (size_t)2 * kTile * 3 * sizeof(float) + (size_t)kBqWarps * nsample * sizeof(int);
  if (smem > 200 * 1024) return EDA_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    EDA_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "ball_query smem attr");
  // bulk copies need 16-byte aligned sources and sizes: every scene/tile starts at a multiple of 4 points
  const int use_tma = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(xyz) & 15u) == 0);
  const float r2 = radius * radius;  // f32 product, ball_query_gpu.cu:27
  dim3 grid((unsigned)((M + kBqWarps - 1) / kBqWarps), (unsigned)B);
  ball_query_kernel<<<grid, kBqThreads, smem, st>>>(new_xyz, xyz, N, M, r2, nsample, use_tma, idx);
  return check_launch("ball_query_kernel");
}
