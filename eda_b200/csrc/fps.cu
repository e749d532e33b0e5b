// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling (pointnet2/_ext_src/src/sampling.cpp:70-91; kernel
// sampling_gpu.cu:74-178).  The reference runs ONE 512-thread block per scene and re-reads
// xyz + the running minima from global memory in every one of the m-1 serial iterations.
//
// Here one thread-block CLUSTER owns a scene.  Every thread keeps its <= P points (xyz and
// running minimum) in registers for the whole kernel; an iteration is
//     P register updates (packed fp32x2 math: FADD2 / FMUL2 / FFMA2, two points per instruction)
//     -> warp arg-max (2x redux.sync) -> CTA arg-max (1 bar.sync, 8 warps)
//     -> one-sided DSMEM exchange of the CTA winners (st.async + mbarrier complete_tx,
//        no cluster barrier) -> every warp picks the cluster winner and its coordinates.
// The winner's xyz travels with its key, so nothing touches L2/HBM inside the loop.  The loop is
// bound by instruction issue and by the reduction latency chain, so CTAs are kept small (256
// threads, many points per thread): the per-warp reduction overhead is paid by 8 warps, not 32.
//
// Bit-exactness.  The reference's result depends on its thread layout: lane t of BS lanes
// scans k = t, t+BS, ... keeping the first strict maximum, then a shared-memory tree keeps
// the LOWER slot on ties.  That is the total order
//     (d2 desc, bitrev_{log2 BS}(k mod BS) asc, k div BS asc)            [BS = opt_n_threads(N)]
// (verified against the literal emulation in oracle/pointnet2_oracle.c and against the
// reference's compiled kernel, tests/golden).  Any decomposition that reduces with this order gives
// the same index.  A thread here visits its points in ascending order of
//     code = bitrev(k mod BS) << 23 | (k div BS)
// so its local strict-'>' scan already honours the order, and cross-thread reduction compares
// (d2 bits, ~code).  Points the reference skips (|p|^2 <= 1e-3, compared in double) and padding get
// a running minimum of -1: fminf keeps it at -1 forever and -1 never beats the initial best of -1.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace eda {
namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Key {
  int hi;       // float bits of d2 (>= 0), or negative when the unit has no candidate
  unsigned lo;  // ~code
};

struct __align__(16) Cand {  // what one CTA tells its peers: 20 payload bytes
  int hi;
  unsigned lo;
  float x, y;
  float z;
  int pad[3];
};

struct __align__(16) WarpCand {  // a warp's winner: key + where its coordinates sit in s_xyz
  int hi;
  unsigned lo;
  int slot;
  int pad;
};

__device__ __forceinline__ unsigned bitrev(unsigned v, int lb) { return lb ? (__brev(v) >> (32 - lb)) : 0u; }

// Warp arg-max under the reference order.  Two redux.sync instead of a 5-step shuffle tree.
__device__ __forceinline__ Key warp_argmax(int hi, unsigned lo) {
  Key k;
  k.hi = __reduce_max_sync(kFull, hi);
  k.lo = __reduce_max_sync(kFull, hi == k.hi ? lo : 0u);
  return k;
}

// Where a thread's jj-th point lives.  G = threads per scene, BS = 1 << lb reference lanes,
// rows = ceil(N / BS).
//   G >= BS : thread g serves lane g mod BS, rows (g div BS) + jj * (G / BS)
//   G <  BS : thread g serves lanes g + u*G (u < BS/G), all rows of one lane before the next, lanes
//             taken in ascending bit-reversed order (u = bitrev_q(jj div rows))
// Either way the thread's codes ascend with jj.
struct Layout {
  int lb, lG, rows;  // log2 BS, log2 G
  // returns false for a padding slot (no such point); L/row are then only a unique-ish placeholder
  __device__ __forceinline__ bool locate(unsigned g, int jj, unsigned &L, unsigned &row) const {
    if (lG >= lb) {
      L = g & ((1u << lb) - 1u);
      row = (g >> lb) + ((unsigned)jj << (lG - lb));
      return row < (unsigned)rows;
    }
    const int q = lb - lG;
    const unsigned up = (unsigned)jj / (unsigned)rows;
    row = (unsigned)jj - up * (unsigned)rows;
    L = g + (bitrev(up & ((1u << q) - 1u), q) << lG);
    return up < (1u << q);
  }
};

// One CTA writes idxs = 0 .. m-1 and publishes every progress milestone the full kernel would have: milestone j
// has its OWN counter word progress[j] (a single shared word would let the scenes of an early wave release a chunk
// that later scenes have not reached yet).
__device__ __forceinline__ void fps_write_identity(int *__restrict__ idxs, int m, int *__restrict__ progress, int every) {
  for (int i = threadIdx.x; i < m; i += blockDim.x) idxs[i] = i;
  if (progress != nullptr) {
    __threadfence();
    __syncthreads();
    const int nmarks = (m + every - 1) / every;
    for (int j = threadIdx.x; j < nmarks; j += blockDim.x) atomicAdd(progress + j, 1);
  }
}

template <int P2, int CL, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_kernel(const float *__restrict__ xyz_all, int N, int m, Layout lay, int *__restrict__ idxs_all,
                   int *__restrict__ progress, int every, const int *__restrict__ not_identity) {
  constexpr int P = 2 * P2;
  // Verified shortcut (eda_fps_identity_check): the answer for this scene is 0, 1, ..., m-1.  Uniform over the
  // cluster, taken before any barrier exists.
  if (not_identity != nullptr && __ldg(not_identity + blockIdx.x / CL) == 0) {
    if (((CL > 1) ? cluster_ctarank() : 0u) == 0u) fps_write_identity(idxs_all + (size_t)(blockIdx.x / CL) * m, m, progress, every);
    return;
  }
  constexpr int kWarps = T / 32;
  extern __shared__ __align__(16) float s_xyz[];  // [P][T][3] copy of this CTA's points, then [P][T] ~codes
  unsigned *s_code = reinterpret_cast<unsigned *>(s_xyz + (size_t)P * T * 3);
  __shared__ WarpCand s_warp[2][kWarps];
  __shared__ Cand s_cta[2][CL];
  __shared__ __align__(8) uint64_t s_bar[2];

  const unsigned rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / CL;
  const float *__restrict__ xyz = xyz_all + (size_t)scene * N * 3;
  int *__restrict__ idxs = idxs_all + (size_t)scene * m;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned g = rank * T + tid;  // thread id within the scene
  const int lb = lay.lb;

  float2 px[P2], py[P2], pz[P2], t[P2];
#pragma unroll
  for (int i = 0; i < P2; ++i) {
    float cx[2], cy[2], cz[2], ct[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int jj = 2 * i + e;
      unsigned L, row;
      const bool slot_ok = lay.locate(g, jj, L, row);
      const long long k = (long long)L + ((long long)row << lb);
      float x = 0.f, y = 0.f, z = 0.f, tt = -1.0f;
      if (slot_ok && k < N) {
        x = __ldg(xyz + k * 3 + 0);
        y = __ldg(xyz + k * 3 + 1);
        z = __ldg(xyz + k * 3 + 2);
        const float mag = sq3(x, y, z);
        tt = ((double)mag <= 1e-3) ? -1.0f : (float)1e10;  // sampling_gpu.cu:105-106, sampling.cpp:78-80
      }
      cx[e] = x; cy[e] = y; cz[e] = z; ct[e] = tt;
      float *s = s_xyz + ((size_t)jj * T + tid) * 3;
      s[0] = x; s[1] = y; s[2] = z;
      s_code[jj * T + tid] = ~((bitrev(L, lb) << 23) | row);
    }
    px[i] = make_float2(cx[0], cx[1]);
    py[i] = make_float2(cy[0], cy[1]);
    pz[i] = make_float2(cz[0], cz[1]);
    t[i] = make_float2(ct[0], ct[1]);
  }
  const float p0x = __ldg(xyz + 0), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float ox = p0x, oy = p0y, oz = p0z;

  if (CL > 1) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_fence_init_cluster();
    }
    cluster_sync_all();  // peers' barriers exist before anyone sends
  } else {
    __syncthreads();
  }
  if (rank == 0 && tid == 0) idxs[0] = 0;

  int next_mark = (progress != nullptr) ? min(every, m) : 0;  // sample count at which the next milestone is published
  int mark = 0;                                               // its index = which counter word it increments
  for (int it = 1; it < m; ++it) {
    const int buf = it & 1;
    if (CL > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[buf], CL * 20);

    // ---- distance update, thread-local arg-max (first strict maximum, ascending code) ----
    // (p - o) as p + (-o): same rounding; two points per FADD2 / FMUL2 / FFMA2
    const float2 nx = make_float2(-ox, -ox), ny = make_float2(-oy, -oy), nz = make_float2(-oz, -oz);
    float val[P];
#pragma unroll
    for (int i = 0; i < P2; ++i) {
      const float2 dx = __fadd2_rn(px[i], nx), dy = __fadd2_rn(py[i], ny), dz = __fadd2_rn(pz[i], nz);
      const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));  // sq3 order, per component
      const float a = fminf(d.x, t[i].x), b = fminf(d.y, t[i].y);
      t[i] = make_float2(a, b);
      val[2 * i] = a;
      val[2 * i + 1] = b;
    }
    // first strict maximum in slot order as a tournament (depth log2 P instead of a P-long dependency
    // chain): the later slot wins only if strictly greater, so ties keep the earlier slot at every level
    int sel[P];
#pragma unroll
    for (int j = 0; j < P; ++j) sel[j] = j;
#pragma unroll
    for (int stride = 1; stride < P; stride *= 2) {
#pragma unroll
      for (int j = 0; j + stride < P; j += 2 * stride) {
        const bool later = val[j + stride] > val[j];
        sel[j] = later ? sel[j + stride] : sel[j];
        val[j] = fmaxf(val[j], val[j + stride]);
      }
    }
    const float best = val[0];
    const int bj = sel[0];
    const int hi = __float_as_int(best);
    const unsigned lo = s_code[bj * T + tid];
    const Key wk = warp_argmax(hi, lo);
    if (hi == wk.hi && lo == wk.lo) {  // at most one lane (codes are unique); none if no candidate
      WarpCand c;
      c.hi = hi; c.lo = lo; c.slot = bj * T + tid; c.pad = 0;
      s_warp[buf][warp] = c;
    }
    __syncthreads();

    Key fk;  // scene-wide winner
    int hi2 = -0x7fffffff;
    unsigned lo2 = 0;
    int slot2 = 0;
    if (lane < kWarps) {
      const WarpCand c = s_warp[buf][lane];
      hi2 = c.hi; lo2 = c.lo; slot2 = c.slot;
    }
    const Key ck = warp_argmax(hi2, lo2);  // this CTA's winner (every warp computes it)
    if (CL > 1) {
      if (warp == 0) {
        // the lane holding the CTA winner fetches its coordinates and broadcasts them inside the warp
        const unsigned owner = __ballot_sync(kFull, hi2 == ck.hi && lo2 == ck.lo && ck.hi >= 0);
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (owner) {
          const int src = __ffs(owner) - 1;
          const int slot = __shfl_sync(kFull, slot2, src);
          const float *s = s_xyz + (size_t)slot * 3;
          cx = s[0]; cy = s[1]; cz = s[2];
        }
        if (lane < CL) {
          const uint32_t dst = mapa_u32(smem_u32(&s_cta[buf][rank]), lane);
          const uint32_t rbar = mapa_u32(smem_u32(&s_bar[buf]), lane);
          st_async_v4(dst, (uint32_t)ck.hi, ck.lo, __float_as_uint(cx), __float_as_uint(cy), rbar);
          st_async_b32(dst + 16, __float_as_uint(cz), rbar);
        }
      }
      mbar_wait_cluster(&s_bar[buf], ((it - 1) >> 1) & 1);
      int hi3 = -0x7fffffff;
      unsigned lo3 = 0;
      if (lane < CL) { hi3 = s_cta[buf][lane].hi; lo3 = s_cta[buf][lane].lo; }
      fk = warp_argmax(hi3, lo3);
      if (fk.hi >= 0) {
        const int w = __ffs(__ballot_sync(kFull, hi3 == fk.hi && lo3 == fk.lo)) - 1;
        ox = s_cta[buf][w].x; oy = s_cta[buf][w].y; oz = s_cta[buf][w].z;
      }
    } else {
      fk = ck;
      if (fk.hi >= 0) {
        const unsigned owner = __ballot_sync(kFull, hi2 == ck.hi && lo2 == ck.lo);
        const int slot = __shfl_sync(kFull, slot2, __ffs(owner) - 1);
        const float *s = s_xyz + (size_t)slot * 3;
        ox = s[0]; oy = s[1]; oz = s[2];
      }
    }
    if (fk.hi < 0) {  // every point skipped: the reference leaves besti = 0 (sampling_gpu.cu:95)
      ox = p0x; oy = p0y; oz = p0z;
    }
    if (rank == 0 && tid == 0) {
      int old = 0;
      if (fk.hi >= 0) {
        const unsigned code = ~fk.lo;
        old = (int)(bitrev(code >> 23, lb) + ((code & 0x7fffffu) << lb));
      }
      idxs[it] = old;
      // progress milestones (eda_furthest_point_sampling_progress): idxs[0..it] of this scene are final and
      // visible device-wide before the counter moves, so a consumer released by a stream-ordered wait on the
      // counter may start on the first (it + 1) centres while the sampling continues
      if (it + 1 == next_mark) {  // (no per-iteration division: thread 0's warp is on the critical chain)
        next_mark = min(next_mark + every, m);
        __threadfence();
        atomicAdd(progress + mark++, 1);
      }
    }
  }
  if (CL > 1) cluster_sync_all();  // nobody leaves while a peer may still address its smem
}

// Latency floor of the sampler's reduction / exchange chain (measurement aid behind eda_selftest_fps_exchange): the
// loop of fps_cluster_kernel with the distance sweep and the per-thread tournament removed — per iteration one warp
// arg-max (2 x redux.sync), the shared-memory stage + bar.sync, the CTA arg-max, the one-sided DSMEM exchange of the
// CL winners (st.async + mbarrier complete_tx), the wake-up and the cluster arg-max.  Each iteration's keys depend
// on the previous winner, so iterations serialise exactly as in the sampler.  cycles / iteration of THIS kernel is
// what no amount of bandwidth or issue slots can remove from the sampler at this decomposition.
template <int CL, int T>
__global__ void __launch_bounds__(T, 1) fps_exchange_floor_kernel(int iters, unsigned *__restrict__ sink) {
  constexpr int kWarps = T / 32;
  __shared__ WarpCand s_warp[2][kWarps];
  __shared__ Cand s_cta[2][CL];
  __shared__ __align__(8) uint64_t s_bar[2];
  const unsigned rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (CL > 1) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_fence_init_cluster();
    }
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  unsigned prev = 0x9E3779B9u * (blockIdx.x / CL + 1);
  for (int it = 1; it < iters; ++it) {
    const int buf = it & 1;
    if (CL > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[buf], CL * 20);
    // a key that depends on the previous winner (one multiply-xor: the dependency, not the work)
    const unsigned h = (prev ^ (rank * T + tid + 1)) * 0x85EBCA77u;
    const int hi = (int)(h >> 2);
    const unsigned lo = ~(rank * T + tid);
    const Key wk = warp_argmax(hi, lo);
    if (hi == wk.hi && lo == wk.lo) {
      WarpCand c;
      c.hi = hi; c.lo = lo; c.slot = tid; c.pad = 0;
      s_warp[buf][warp] = c;
    }
    __syncthreads();
    int hi2 = -0x7fffffff;
    unsigned lo2 = 0;
    if (lane < kWarps) { hi2 = s_warp[buf][lane].hi; lo2 = s_warp[buf][lane].lo; }
    const Key ck = warp_argmax(hi2, lo2);
    Key fk = ck;
    if (CL > 1) {
      if (warp == 0 && lane < CL) {
        const uint32_t dst = mapa_u32(smem_u32(&s_cta[buf][rank]), lane);
        const uint32_t rbar = mapa_u32(smem_u32(&s_bar[buf]), lane);
        st_async_v4(dst, (uint32_t)ck.hi, ck.lo, 0u, 0u, rbar);
        st_async_b32(dst + 16, 0u, rbar);
      }
      mbar_wait_cluster(&s_bar[buf], ((it - 1) >> 1) & 1);
      int hi3 = -0x7fffffff;
      unsigned lo3 = 0;
      if (lane < CL) { hi3 = s_cta[buf][lane].hi; lo3 = s_cta[buf][lane].lo; }
      fk = warp_argmax(hi3, lo3);
    }
    prev = (unsigned)fk.hi ^ fk.lo;
  }
  if (tid == 0 && rank == 0) sink[blockIdx.x / CL] = prev;
  if (CL > 1) cluster_sync_all();
}

template <int CL, int T>
int launch_exchange_floor(int B, int iters, unsigned *sink, cudaStream_t st) {
  auto kern = fps_exchange_floor_kernel<CL, T>;
  if (CL > 8) EDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "floor cluster attr");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CL));
  cfg.blockDim = dim3(T);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EDA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, iters, sink), "fps_exchange_floor_kernel launch");
  return check_launch("fps_exchange_floor_kernel");
}

// Generic fallback: any N, running minima in global scratch (the reference's layout), one
// 1024-thread CTA per scene, same ordering rule.  Used only when the register variant does
// not cover the shape or a cluster launch is not possible.
constexpr int kGThreads = 1024;
__global__ void __launch_bounds__(kGThreads, 1)
fps_global_kernel(const float *__restrict__ xyz_all, int N, int m, int lb, float *__restrict__ temp_all,
                  int *__restrict__ idxs_all, int *__restrict__ progress, int every, const int *__restrict__ not_identity) {
  __shared__ Key s_warp[2][kGThreads / 32];
  const int scene = blockIdx.x;
  if (not_identity != nullptr && __ldg(not_identity + scene) == 0) {
    fps_write_identity(idxs_all + (size_t)scene * m, m, progress, every);
    return;
  }
  const float *__restrict__ xyz = xyz_all + (size_t)scene * N * 3;
  float *__restrict__ temp = temp_all + (size_t)scene * N;
  int *__restrict__ idxs = idxs_all + (size_t)scene * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned L = tid & ((1u << lb) - 1u);
  const unsigned rg = tid >> lb;
  const unsigned RG = kGThreads >> lb;
  const unsigned codeL = bitrev(L, lb) << 23;
  const long long rows = ((long long)N + (1ll << lb) - 1) >> lb;

  for (long long r = rg; r < rows; r += RG) {
    const long long k = (long long)L + (r << lb);
    if (k < N) {
      const float mag = sq3(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2]);
      temp[k] = ((double)mag <= 1e-3) ? -1.0f : (float)1e10;
    }
  }
  const float p0x = xyz[0], p0y = xyz[1], p0z = xyz[2];
  float ox = p0x, oy = p0y, oz = p0z;
  if (tid == 0) idxs[0] = 0;
  __syncthreads();
  int next_mark = (progress != nullptr) ? min(every, m) : 0;
  int mark = 0;
  for (int it = 1; it < m; ++it) {
    const int buf = it & 1;
    float best = -1.0f;
    unsigned brow = 0;
    for (long long r = rg; r < rows; r += RG) {
      const long long k = (long long)L + (r << lb);
      if (k < N) {
        const float d = sq3(__fsub_rn(xyz[k * 3], ox), __fsub_rn(xyz[k * 3 + 1], oy), __fsub_rn(xyz[k * 3 + 2], oz));
        const float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > best) { best = d2; brow = (unsigned)r; }
      }
    }
    const Key wk = warp_argmax(__float_as_int(best), ~(codeL | brow));
    if (lane == 0) s_warp[buf][warp] = wk;
    __syncthreads();
    const Key mine = s_warp[buf][lane];
    const Key fk = warp_argmax(mine.hi, mine.lo);
    int old = 0;
    if (fk.hi >= 0) {
      const unsigned code = ~fk.lo;
      old = (int)(bitrev(code >> 23, lb) + ((code & 0x7fffffu) << lb));
    }
    ox = xyz[(size_t)old * 3]; oy = xyz[(size_t)old * 3 + 1]; oz = xyz[(size_t)old * 3 + 2];
    if (tid == 0) {
      idxs[it] = old;
      if (it + 1 == next_mark) {  // (no per-iteration division: thread 0's warp is on the critical chain)
        next_mark = min(next_mark + every, m);
        __threadfence();
        atomicAdd(progress + mark++, 1);
      }
    }
  }
}

// include/cuda_utils.h:18-22 of the reference, evaluated with the same libm expression.
int ref_log2_block(int n) {
  const int pow_2 = (int)(log((double)n) / log(2.0));
  int lb = pow_2 < 0 ? 0 : pow_2;
  if (lb > 9) lb = 9;
  return lb;
}

struct FpsPlan {
  int cl;  // cluster size, 0 = global fallback
  int t;   // threads per CTA
  int p2;  // point PAIRS per thread (template instance)
  Layout lay;
};

constexpr int kP2s[] = {1, 2, 4, 6, 8, 10, 13};
constexpr int kMaxP2 = 13;

int round_p2(int need_points) {
  for (int p2 : kP2s)
    if (need_points <= 2 * p2) return p2;
  return 0;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << (l + 1)) <= v) ++l;
  return l;
}

// points per thread when G threads share one scene
int need_points(int N, int lb, int G) {
  const int BS = 1 << lb;
  const long long rows = ((long long)N + BS - 1) / BS;
  if (G >= BS) return (int)((rows + (G / BS) - 1) / (G / BS));
  return (int)(rows * (BS / G));
}

bool try_plan(int N, int lb, int cl, int t, FpsPlan *out) {
  const int G = cl * t;
  const int need = need_points(N, lb, G);
  const int p2 = round_p2(need);
  if (!p2) return false;
  if (t > 512 && p2 > 4) return false;  // 1024 threads: <= 64 registers per thread
  out->cl = cl; out->t = t; out->p2 = p2;
  out->lay.lb = lb; out->lay.lG = ilog2(G);
  out->lay.rows = (int)(((long long)N + (1 << lb) - 1) >> lb);
  return true;
}

FpsPlan plan_fps(int B, int N, int lb) {
  (void)B;
  FpsPlan pl = {};
  // EDA_FPS_CLUSTER / EDA_FPS_THREADS force a decomposition (tests sweep them: the result may not change)
  const char *e = getenv("EDA_FPS_CLUSTER");
  const char *et = getenv("EDA_FPS_THREADS");
  const int force = e ? atoi(e) : 0;
  const int force_t = et ? atoi(et) : 0;
  if (force == 1 || force == 2 || force == 4 || force == 8 || force == 16) {
    for (int t : {256, 512, 1024})
      if ((force_t == 0 || force_t == t) && try_plan(N, lb, force, t, &pl)) return pl;
  }
  // Smallest cluster whose threads can hold the scene in registers: the serial chain is latency/issue
  // bound, extra CTAs only help by shrinking the per-thread sweep.  16-CTA clusters fit one per GPC, so
  // 8 scenes would run in two waves: last resort.
  for (int cl : {1, 2, 4, 8})
    if (try_plan(N, lb, cl, 256, &pl)) return pl;
  for (int cl : {8, 16})
    for (int t : {256, 512})
      if (try_plan(N, lb, cl, t, &pl)) return pl;
  pl.cl = 0;
  return pl;
}

template <int P2, int CL, int T>
int launch_cluster(const float *xyz, int B, int N, int m, const Layout &lay, int *idxs, int *progress, int every,
                   const int *not_identity,
                   cudaStream_t st) {
  auto kern = fps_cluster_kernel<P2, CL, T>;
  const size_t smem = (size_t)2 * P2 * T * 4 * sizeof(float);
  static SmemAttr smem_attr;  // per template instance
  static std::atomic<unsigned long long> nonportable_devs{0};
  EDA_CUDA_TRY(smem_attr.ensure(kern, smem), "fps smem attr");
  if (CL > 8) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!bit || !(nonportable_devs.load(std::memory_order_acquire) & bit)) {
      EDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "fps cluster attr");
      nonportable_devs.fetch_or(bit, std::memory_order_release);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CL));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EDA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, xyz, N, m, lay, idxs, progress, every, not_identity), "fps_cluster_kernel launch");
  return check_launch("fps_cluster_kernel");
}

template <int CL, int T>
int dispatch_p(const FpsPlan &pl, const float *xyz, int B, int N, int m, int *idxs, int *progress, int every,
               const int *not_identity,
               cudaStream_t st) {
  if constexpr (T == 1024) {
    switch (pl.p2) {
      case 1: return launch_cluster<1, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
      case 2: return launch_cluster<2, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
      case 4: return launch_cluster<4, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    }
    return EDA_ERR_UNSUPPORTED;
  }
  switch (pl.p2) {
    case 1: return launch_cluster<1, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 2: return launch_cluster<2, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 4: return launch_cluster<4, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 6: return launch_cluster<6, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 8: return launch_cluster<8, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 10: return launch_cluster<10, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
    case 13: return launch_cluster<13, CL, T>(xyz, B, N, m, pl.lay, idxs, progress, every, not_identity, st);
  }
  return EDA_ERR_UNSUPPORTED;
}

template <int CL>
int dispatch_t(const FpsPlan &pl, const float *xyz, int B, int N, int m, int *idxs, int *progress, int every,
               const int *not_identity,
               cudaStream_t st) {
  if (pl.t == 256) return dispatch_p<CL, 256>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
  if (pl.t == 512) return dispatch_p<CL, 512>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
  if (pl.t == 1024) return dispatch_p<CL, 1024>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
  return EDA_ERR_UNSUPPORTED;
}



// ---- identity verification for FPS of an FPS-ordered set -----------------------------------------------------------
// The backbone samples 2048 -> 1024 -> 512 -> 256 from sets that are already in FPS order: without
// exact ties the answer is 0, 1, ..., m-1, and 1791 strictly serial iterations (0.63 ms at B = 8) reproduce it.  The
// answer IS the identity iff at every step i < m point i is the strict maximiser of the running minimum, i.e.
//     d_i := min_{k<i} |p_i - p_k|^2  >  min_{k<i} |p_j - p_k|^2   for every j != i          (and p_i is not skipped)
// (strict, so the reference's thread-layout tie-break never decides; j < i gives d_i > 0).  That condition is
// embarrassingly parallel: pass 1 computes d_i (thread = i), pass 2 checks every j against d_1 .. d_{min(j,m)-1}
// (thread = j, one running minimum), both with the reference's exact distance arithmetic.  A scene that fails any
// comparison (ties, duplicates, skipped points, NaNs) sets its flag and the sampler runs the full algorithm for it.
constexpr int kIdThreads = 256;

__global__ void __launch_bounds__(kIdThreads)
fps_identity_pass1_kernel(const float *__restrict__ xyz_all, int n, int m, float *__restrict__ dsel_all,
                          int *__restrict__ not_identity) {
  extern __shared__ float s_pts[];  // [min(n, m)][3]: only points k < m are ever reference points
  const int scene = blockIdx.y;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  const int np = min(n, m);
  for (int i = threadIdx.x; i < np * 3; i += kIdThreads) s_pts[i] = __ldg(xyz + i);
  __syncthreads();
  const int i = blockIdx.x * kIdThreads + threadIdx.x;
  if (i >= m) return;
  if (i >= n) {  // more samples than points: never the identity
    atomicExch(not_identity + scene, 1);
    return;
  }
  const float x = s_pts[i * 3], y = s_pts[i * 3 + 1], z = s_pts[i * 3 + 2];
  float run = 1e10f;  // the reference's initial running minimum (sampling.cpp:78-80)
  for (int k = 0; k < i; ++k) {
    const float d = sq3(__fadd_rn(x, -s_pts[k * 3]), __fadd_rn(y, -s_pts[k * 3 + 1]), __fadd_rn(z, -s_pts[k * 3 + 2]));
    run = fminf(run, d);
  }
  dsel_all[(size_t)scene * m + i] = run;
  if (i > 0) {
    const bool skipped = (double)sq3(x, y, z) <= 1e-3;  // sampling_gpu.cu:105-106
    if (skipped || !(run > 0.f)) atomicExch(not_identity + scene, 1);
  }
}

__global__ void __launch_bounds__(kIdThreads)
fps_identity_pass2_kernel(const float *__restrict__ xyz_all, int n, int m, const float *__restrict__ dsel_all,
                          int *__restrict__ not_identity) {
  extern __shared__ float s_pts[];  // [min(n, m)][3] reference points, then [m] d_i
  const int scene = blockIdx.y;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  const int np = min(n, m);
  float *s_d = s_pts + (size_t)np * 3;
  for (int i = threadIdx.x; i < np * 3; i += kIdThreads) s_pts[i] = __ldg(xyz + i);
  for (int i = threadIdx.x; i < np; i += kIdThreads) s_d[i] = __ldg(dsel_all + (size_t)scene * m + i);
  __syncthreads();
  const int j = blockIdx.x * kIdThreads + threadIdx.x;
  if (j >= n) return;
  const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
  // j competes with i = 1 .. min(j, m) - 1 (for i >= j it is the candidate itself or already selected)
  const int last = min(j, np) - 1;
  float run = 1e10f;
  bool bad = false;
  for (int k = 0; k < last; ++k) {
    const float d = sq3(__fadd_rn(x, -s_pts[k * 3]), __fadd_rn(y, -s_pts[k * 3 + 1]), __fadd_rn(z, -s_pts[k * 3 + 2]));
    run = fminf(run, d);               // = min_{k' <= k} |p_j - p_k'|^2 : j's running minimum when step k + 1 is decided
    bad |= !(run < s_d[k + 1]);
  }
  if (bad) atomicExch(not_identity + scene, 1);
}

}  // namespace
}  // namespace eda

extern "C" {

size_t eda_fps_scratch_bytes(int B, int N, int m) {
  if (B <= 0 || N <= 0 || m <= 0) return 0;
  const int lb = eda::ref_log2_block(N);
  const eda::FpsPlan pl = eda::plan_fps(B, N, lb);
  return pl.cl == 0 ? (size_t)B * N * sizeof(float) : 0;
}

int eda_fps_plan(int B, int N, int m, int *cluster, int *threads, int *points_per_thread) {
  (void)m;
  if (B <= 0 || N <= 0) return EDA_ERR_INVALID_ARGUMENT;
  const eda::FpsPlan pl = eda::plan_fps(B, N, eda::ref_log2_block(N));
  if (cluster) *cluster = pl.cl;            // 0: the global-scratch fallback kernel
  if (threads) *threads = pl.cl ? pl.t : eda::kGThreads;
  if (points_per_thread) *points_per_thread = pl.cl ? 2 * pl.p2 : 0;
  return EDA_OK;
}

static int fps_impl(const float *xyz, int B, int N, int m, void *scratch, int *idxs, int *progress, int every,
                    const int *not_identity,
                    void *stream) {
  using namespace eda;
  if (B < 0 || N < 0 || m < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || m == 0) return EDA_OK;
  if (!xyz || !idxs || N == 0) return EDA_ERR_INVALID_ARGUMENT;
  if (progress && every < 2) return EDA_ERR_INVALID_ARGUMENT;
  if (((long long)N >> 9) >= (1 << 23)) return EDA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int lb = ref_log2_block(N);
  const FpsPlan pl = plan_fps(B, N, lb);
  switch (pl.cl) {
    case 1: return dispatch_t<1>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
    case 2: return dispatch_t<2>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
    case 4: return dispatch_t<4>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
    case 8: return dispatch_t<8>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
    case 16: return dispatch_t<16>(pl, xyz, B, N, m, idxs, progress, every, not_identity, st);
    default: break;
  }
  if (!scratch) return EDA_ERR_INVALID_ARGUMENT;
  fps_global_kernel<<<B, kGThreads, 0, st>>>(xyz, N, m, lb, (float *)scratch, idxs, progress, every, not_identity);
  return check_launch("fps_global_kernel");
}

int eda_furthest_point_sampling(const float *xyz, int B, int N, int m, void *scratch, int *idxs, void *stream) {
  return fps_impl(xyz, B, N, m, scratch, idxs, nullptr, 0, nullptr, stream);
}

int eda_furthest_point_sampling_ex(const float *xyz, int B, int N, int m, void *scratch, int *idxs, int *progress,
                                   int every, const int *not_identity, void *stream) {
  return fps_impl(xyz, B, N, m, scratch, idxs, progress, progress ? every : 0, not_identity, stream);
}

int eda_fps_identity_check(const float *xyz, int B, int n, int m, float *dsel, int *not_identity, void *stream) {
  using namespace eda;
  if (B < 0 || n < 0 || m < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || m == 0) return EDA_OK;
  if (!xyz || !dsel || !not_identity || n == 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B > 65535) return EDA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  EDA_CUDA_TRY(cudaMemsetAsync(not_identity, 0, (size_t)B * sizeof(int), st), "identity flag memset");
  const int np = n < m ? n : m;
  const size_t smem1 = (size_t)np * 3 * sizeof(float), smem2 = (size_t)np * 4 * sizeof(float);
  if (smem2 > 200 * 1024) return EDA_ERR_UNSUPPORTED;
  static SmemAttr attr1, attr2;
  if (smem2 > 48 * 1024) {
    EDA_CUDA_TRY(attr1.ensure(fps_identity_pass1_kernel, smem2), "identity smem attr");
    EDA_CUDA_TRY(attr2.ensure(fps_identity_pass2_kernel, smem2), "identity smem attr");
  }
  dim3 g1((unsigned)((m + kIdThreads - 1) / kIdThreads), (unsigned)B), g2((unsigned)((n + kIdThreads - 1) / kIdThreads), (unsigned)B);
  fps_identity_pass1_kernel<<<g1, kIdThreads, smem1, st>>>(xyz, n, m, dsel, not_identity);
  fps_identity_pass2_kernel<<<g2, kIdThreads, smem2, st>>>(xyz, n, m, dsel, not_identity);
  return check_launch("fps_identity_kernels", 2);
}

int eda_selftest_fps_exchange(int B, int cluster, int threads, int iters, unsigned int *sink, void *stream) {
  using namespace eda;
  if (B <= 0 || iters < 2 || !sink) return EDA_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  if (threads == 256) {
    switch (cluster) {
      case 1: return launch_exchange_floor<1, 256>(B, iters, sink, st);
      case 2: return launch_exchange_floor<2, 256>(B, iters, sink, st);
      case 4: return launch_exchange_floor<4, 256>(B, iters, sink, st);
      case 8: return launch_exchange_floor<8, 256>(B, iters, sink, st);
      case 16: return launch_exchange_floor<16, 256>(B, iters, sink, st);
    }
  } else if (threads == 512) {
    switch (cluster) {
      case 4: return launch_exchange_floor<4, 512>(B, iters, sink, st);
      case 8: return launch_exchange_floor<8, 512>(B, iters, sink, st);
      case 16: return launch_exchange_floor<16, 512>(B, iters, sink, st);
    }
  }
  return EDA_ERR_UNSUPPORTED;
}

int eda_furthest_point_sampling_progress(const float *xyz, int B, int N, int m, void *scratch, int *idxs,
                                         int *progress, int every, void *stream) {
  if (!progress) return EDA_ERR_INVALID_ARGUMENT;
  return fps_impl(xyz, B, N, m, scratch, idxs, progress, every, nullptr, stream);
}

// Stream-ordered wait on a device word: work queued on `stream` after this call starts once
// (int)(*addr - value) >= 0.  No SM is occupied while waiting (cuStreamWaitValue32, resolved through the runtime
// so the library does not link against libcuda).
int eda_stream_wait_value32(void *stream, const int *addr, int value) {
  using namespace eda;
  if (!addr) return EDA_ERR_INVALID_ARGUMENT;
  typedef int (*wait_fn_t)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
  static wait_fn_t fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    EDA_CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &qres), "cuStreamWaitValue32 lookup");
    if (!sym || qres != cudaDriverEntryPointSuccess) return EDA_ERR_UNSUPPORTED;
    fn = reinterpret_cast<wait_fn_t>(sym);
  }
  const int rc = fn(as_stream(stream), (unsigned long long)reinterpret_cast<uintptr_t>(addr), (unsigned int)value,
                    0u /* CU_STREAM_WAIT_VALUE_GEQ */);
  if (rc != 0) {
    set_last_cuda_error(cudaErrorUnknown, "cuStreamWaitValue32");
    return EDA_ERR_CUDA_LAUNCH;
  }
  return EDA_OK;
}

}  // extern "C"
