// Ball query for sm_100a.
//
// Replaces ball_query (pointnet2/_ext_src/src/ball_query.cpp:13-37; kernel
// ball_query_gpu.cu:14-49): for every centre, the first `nsample` point indices in ascending
// index order with d2 < radius^2 (strict, f32 radius*radius), remaining slots padded with the
// first hit, all zeros for an empty ball.
//
// The reference gives each centre to ONE thread of ONE block per scene and walks xyz in
// global memory.  Here a CTA owns kBqWarps * kQ centres and streams the scene's xyz through shared
// memory in tiles fetched by the TMA engine (1-D cp.async.bulk, double buffered, mbarrier
// completion).  The kernel is bound by instruction issue, so the inner loop is register-tiled: a
// lane loads a point once and tests it against the warp's kQ = 4 centres with packed fp32x2 math
// (two centres per FADD2 / FMUL2 / FFMA2), 64 points per warp step.  Hits are rare, so the common
// step is branch-free; the ordered compaction (ballot + prefix popcount) keeps "first nsample in
// index order" exact.  Rows are staged in smem and written once, coalesced, including the padding —
// idx needs no zero-fill by the caller.
#include "common.cuh"

namespace eda {
namespace {

constexpr int kBqWarps = 8;
constexpr int kBqThreads = kBqWarps * 32;
constexpr int kQ = 4;        // centres per warp
constexpr int kTile = 2048;  // points per smem tile: 24 KB, x2 buffers
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float *__restrict__ new_xyz_all, const int *__restrict__ centre_idx,
                  const float *__restrict__ xyz_all, int N, int M, int Mtot, int m0, float r2, int nsample, int use_tma,
                  int *__restrict__ idx_all) {
  // Centres m0 .. m0+M-1 of the Mtot centres of every scene (the plain op has m0 = 0, M = Mtot).  With
  // centre_idx the centre coordinates are read through the FPS indices (xyz[centre_idx]) instead of new_xyz:
  // bit-identical, and it lets the query start before the whole FPS result (and its gather) exists.
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile0 = reinterpret_cast<float *>(smem_raw);
  float *tile1 = tile0 + kTile * 3;
  int *hits = reinterpret_cast<int *>(tile1 + kTile * 3);  // [kBqWarps][kQ][nsample]
  __shared__ __align__(8) uint64_t full[2];

  const int b = blockIdx.y;
  const float *__restrict__ xyz = xyz_all + (size_t)b * N * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int centre0 = (blockIdx.x * kBqWarps + warp) * kQ;
  int *myhits = hits + warp * kQ * nsample;

  // centre coordinates as pairs: (c0,c1) and (c2,c3); centres past M start "already full", their hits are ignored
  float2 cx[2], cy[2], cz[2];
  int cnt[kQ];
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const int c = centre0 + q;
    float x = 0.f, y = 0.f, z = 0.f;
    if (c < M) {
      const size_t ci = (size_t)b * Mtot + m0 + c;
      const float *p = centre_idx ? xyz + (size_t)__ldg(centre_idx + ci) * 3 : new_xyz_all + ci * 3;
      x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
    if (q & 1) { cx[q >> 1].y = x; cy[q >> 1].y = y; cz[q >> 1].y = z; }
    else       { cx[q >> 1].x = x; cy[q >> 1].x = y; cz[q >> 1].x = z; }
    cnt[q] = (c < M) ? 0 : nsample;  // "already full": its hits are ignored
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

  auto all_full = [&]() {
    bool f = true;
#pragma unroll
    for (int q = 0; q < kQ; ++q) f = f && (cnt[q] >= nsample);
    return f;
  };
  bool done = all_full();
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
      for (int base = 0; base < npts && !done; base += 64) {
        bool hit[2][kQ];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = base + u * 32 + lane;
          const int kk = min(k, npts - 1);
          const float x = tile[kk * 3 + 0], y = tile[kk * 3 + 1], z = tile[kk * 3 + 2];
          const float2 mx = make_float2(-x, -x), my = make_float2(-y, -y), mz = make_float2(-z, -z);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // (new_x - x)^2 + ... as the reference compiles it: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
            const float2 dx = __fadd2_rn(cx[h], mx), dy = __fadd2_rn(cy[h], my), dz = __fadd2_rn(cz[h], mz);
            const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
            hit[u][2 * h + 0] = (k < npts) && (d2.x < r2);
            hit[u][2 * h + 1] = (k < npts) && (d2.y < r2);
          }
        }
        bool any = false;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < kQ; ++q) any = any || hit[u][q];
        if (__any_sync(kFull, any)) {
#pragma unroll
          for (int q = 0; q < kQ; ++q) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const unsigned mask = __ballot_sync(kFull, hit[u][q]);
              if (mask && cnt[q] < nsample) {
                if (hit[u][q]) {
                  const int pos = cnt[q] + __popc(mask & ((1u << lane) - 1u));
                  if (pos < nsample) myhits[q * nsample + pos] = kbase + base + u * 32 + lane;
                }
                cnt[q] = min(nsample, cnt[q] + __popc(mask));
              }
            }
          }
          done = all_full();
        }
      }
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

  __syncwarp();
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const int c = centre0 + q;
    if (c < M) {
      int *out = idx_all + ((size_t)b * Mtot + m0 + c) * nsample;
      const int have = cnt[q];
      const int first = have > 0 ? myhits[q * nsample] : 0;  // ball_query_gpu.cu:38-42 pads with the first hit
      for (int s = lane; s < nsample; s += 32) out[s] = s < have ? myhits[q * nsample + s] : first;
    }
  }
}

}  // namespace
}  // namespace eda

namespace eda {
namespace {
int launch_ball_query(const float *new_xyz, const int *centre_idx, const float *xyz, int B, int N, int M, int Mtot, int m0,
                      float radius, int nsample, int *idx, cudaStream_t st) {
  if (B < 0 || N < 0 || M < 0 || nsample < 0 || m0 < 0 || m0 + M > Mtot) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0 || nsample == 0) return EDA_OK;
  if ((!new_xyz && !centre_idx) || !idx || (N > 0 && !xyz)) return EDA_ERR_INVALID_ARGUMENT;
  if (N == 0) {  // nothing to find: the reference returns its zero-initialised tensor
    if (M != Mtot) return EDA_ERR_UNSUPPORTED;
    EDA_CUDA_TRY(cudaMemsetAsync(idx, 0, (size_t)B * M * nsample * sizeof(int), st), "ball_query memset");
    return EDA_OK;
  }
  const size_t smem = (size_t)2 * kTile * 3 * sizeof(float) + (size_t)kBqWarps * kQ * nsample * sizeof(int);
  if (smem > 200 * 1024) return EDA_ERR_UNSUPPORTED;
  static SmemAttr attr;
  if (smem > 48 * 1024) EDA_CUDA_TRY(attr.ensure(ball_query_kernel, smem), "ball_query smem attr");
  // bulk copies need 16-byte aligned sources and sizes: every scene/tile starts at a multiple of 4 points
  const int use_tma = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(xyz) & 15u) == 0);
  const float r2 = radius * radius;  // f32 product, ball_query_gpu.cu:27
  dim3 grid((unsigned)((M + kBqWarps * kQ - 1) / (kBqWarps * kQ)), (unsigned)B);
  ball_query_kernel<<<grid, kBqThreads, smem, st>>>(new_xyz, centre_idx, xyz, N, M, Mtot, m0, r2, nsample, use_tma, idx);
  return check_launch("ball_query_kernel");
}
}  // namespace
}  // namespace eda

extern "C" int eda_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius, int nsample,
                              int *idx, void *stream) {
  if (!new_xyz && B > 0 && M > 0 && nsample > 0) return EDA_ERR_INVALID_ARGUMENT;
  return eda::launch_ball_query(new_xyz, nullptr, xyz, B, N, M, M, 0, radius, nsample, idx, eda::as_stream(stream));
}

extern "C" int eda_ball_query_range(const float *xyz, const int *centre_idx, int B, int N, int Mtot, int m0, int mc,
                                    float radius, int nsample, int *idx, void *stream) {
  if (!centre_idx && B > 0 && mc > 0 && nsample > 0) return EDA_ERR_INVALID_ARGUMENT;
  return eda::launch_ball_query(nullptr, centre_idx, xyz, B, N, mc, Mtot, m0, radius, nsample, idx,
                                eda::as_stream(stream));
}
