// Backward building blocks shared by the attention layers and the set-abstraction MLP (sm_100a):
//
//   eda_wgrad               dW[N x K] += dY[R x N]^T X[R x K]  (+ db[N] += column sums of dY), up to 6 problems
//                           of one (N, K) per launch, split over the row dimension R.  This is the weight
//                           gradient of every nn.Linear / 1x1 conv on the path (autograd's addmm / mm backward
//                           in the reference: one cuBLAS GEMM + one reduction each).  The contraction index is
//                           the ROW of both operands, i.e. both are "MN-major" for the tensor core; they are
//                           consumed as they lie in memory by warp-level mma.sync.m16n8k8 tf32 fragments read
//                           from a row-major shared-memory tile (no transposition pass).
//   eda_layernorm_backward  du = LayerNorm'(u) dy, dgamma += sum dy * xhat, dbeta += sum dy, and the dropout mask
//                           of the block output re-applied (the same counter-based hash as the forward GEMM
//                           epilogue), warp = row.
//   eda_relu_backward       dz = dy * [y > 0] * scale   (FFN hidden layer; y is the saved post-ReLU activation)
//
// The activation-gradient GEMMs (dX = dY W) reuse the forward tcgen05 kernel (eda_linear_forward) with the
// transposed weight packed by eda_linear_pack_strided.
#include "common.cuh"
#include "wgrad_tc.h"

namespace eda {
namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------
constexpr int kWgMaxProbs = 6;
constexpr int kWgTile = 64;    // output tile: 64 (n, columns of dY) x 64 (k, columns of X)
constexpr int kWgRows = 32;    // rows per staged chunk
constexpr int kWgPitch = 72;   // floats per shared-memory row: (t * 72 + g) % 32 = 8 t + g -> conflict-free fragments
constexpr int kWgThreads = 128;

struct WgProblem {
  const float *dy, *x;
  const float *x_scale, *x_shift;  // optional: x is consumed as relu(x * x_scale[k] + x_shift[k])
  float *dw, *db;
  long long rows;
  int ldy, ldx, ldw;
};
struct WgParams {
  WgProblem pr[kWgMaxProbs];
  int nprobs, N, K, splits;
};

__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(const WgParams p) {
  __shared__ __align__(16) float sY[2][kWgRows][kWgPitch];
  __shared__ __align__(16) float sX[2][kWgRows][kWgPitch];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int pi = (int)blockIdx.z / p.splits, split = (int)blockIdx.z - pi * p.splits;
  const WgProblem &pr = p.pr[pi];
  const int n0 = blockIdx.x * kWgTile, k0 = blockIdx.y * kWgTile;
  // rows of this split, in whole chunks
  const long long chunks_total = (pr.rows + kWgRows - 1) / kWgRows;
  const long long per = (chunks_total + p.splits - 1) / p.splits;
  const long long c_lo = (long long)split * per;
  const long long c_hi = c_lo + per < chunks_total ? c_lo + per : chunks_total;
  if (c_lo >= c_hi) return;
  const bool want_db = pr.db != nullptr && blockIdx.y == 0;

  // staging: 32 rows x 16 float4 per matrix = 512 chunks, 4 per thread
  auto issue = [&](long long c, int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = i * kWgThreads + tid;
      const int r = id >> 4, c4 = (id & 15) << 2;
      const long long row = c * kWgRows + r;
      const bool rin = row < pr.rows;
      const bool yin = rin && n0 + c4 < p.N;
      const bool xin = rin && k0 + c4 < p.K;
      cp_async16_zfill(&sY[buf][r][c4], yin ? pr.dy + row * pr.ldy + n0 + c4 : pr.dy, yin ? 16u : 0u);
      cp_async16_zfill(&sX[buf][r][c4], xin ? pr.x + row * pr.ldx + k0 + c4 : pr.x, xin ? 16u : 0u);
    }
    cp_async_commit_group();
  };

  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;  // warp tile 32 (n) x 32 (k)
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  float bsum = 0.f;
  // optional prologue on X (folded BatchNorm + ReLU of the producing layer), per column of this thread's B fragments
  const bool pro = pr.x_scale != nullptr;
  float xs[4], xh[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int k = k0 + wn + nt * 8 + g;
    xs[nt] = (pro && k < p.K) ? __ldg(pr.x_scale + k) : 1.f;
    xh[nt] = (pro && k < p.K) ? __ldg(pr.x_shift + k) : 0.f;
  }

  issue(c_lo, 0);
  for (long long c = c_lo; c < c_hi; ++c) {
    const int buf = (int)((c - c_lo) & 1);
    if (c + 1 < c_hi) {
      issue(c + 1, buf ^ 1);
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kWgRows / 8; ++ks) {
      uint32_t a[2][4], b[4][2];
      const float *y0 = &sY[buf][ks * 8 + t][0], *y1 = &sY[buf][ks * 8 + t + 4][0];
      const float *x0 = &sX[buf][ks * 8 + t][0], *x1 = &sX[buf][ks * 8 + t + 4][0];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int m = wm + mt * 16 + g;
        a[mt][0] = f2tf32(y0[m]);
        a[mt][1] = f2tf32(y0[m + 8]);
        a[mt][2] = f2tf32(y1[m]);
        a[mt][3] = f2tf32(y1[m + 8]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn + nt * 8 + g;
        float v0 = x0[n], v1 = x1[n];
        if (pro) {
          v0 = fmaxf(fmaf(v0, xs[nt], xh[nt]), 0.f);
          v1 = fmaxf(fmaf(v1, xs[nt], xh[nt]), 0.f);
        }
        b[nt][0] = f2tf32(v0);
        b[nt][1] = f2tf32(v1);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_m16n8k8_tf32(acc[mt][nt], a[mt], b[nt]);
    }
    if (want_db && tid < kWgTile) {
#pragma unroll 8
      for (int r = 0; r < kWgRows; ++r) bsum += sY[buf][r][tid];
    }
    __syncthreads();  // the buffer is refilled by the next iteration's prefetch
  }

  // accumulator element (mt, nt, e): n = wm + mt*16 + g + (e >= 2 ? 8 : 0), k = wn + nt*8 + 2t + (e & 1)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + wm + mt * 16 + g + ((e & 2) ? 8 : 0);
        const int k = k0 + wn + nt * 8 + 2 * t + (e & 1);
        if (n < p.N && k < p.K) atomicAdd(pr.dw + (size_t)n * pr.ldw + k, acc[mt][nt][e]);
      }
  if (want_db && tid < kWgTile && n0 + tid < p.N) atomicAdd(pr.db + n0 + tid, bsum);
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward (+ output-dropout mask), warp = row
// ---------------------------------------------------------------------------------------------------------
constexpr int kLnMaxN = 384;
constexpr int kLnJ = kLnMaxN / 128;  // float4 columns per lane
constexpr int kLnWarps = 8;

__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_backward_kernel(const float *__restrict__ dy, const float *__restrict__ u, const float *__restrict__ gamma,
                          float eps, long long rows, int N, float *__restrict__ du, float *__restrict__ dproj,
                          float *__restrict__ dgamma, float *__restrict__ dbeta, uint32_t drop_thresh,
                          uint32_t drop_seed_base, float drop_scale, const uint32_t *seed_epoch) {
  pdl_launch_dependents();  // a PDL-launched successor may be scheduled now (it waits for this grid's completion itself)
  pdl_wait();               // launched with the PDL attribute: dy comes from the kernel just before
  __shared__ float s_acc[2][kLnMaxN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n4 = N >> 2;
  const uint32_t drop_seed = dproj ? effective_seed(drop_seed_base, seed_epoch) : 0u;
  for (int i = tid; i < 2 * kLnMaxN; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  float4 g4[kLnJ], ag[kLnJ], ab[kLnJ];
#pragma unroll
  for (int j = 0; j < kLnJ; ++j) {
    const int c4 = lane + j * 32;
    g4[j] = (c4 < n4 && gamma) ? __ldg(reinterpret_cast<const float4 *>(gamma) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
    ag[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invN = 1.0f / (float)N;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    float4 x[kLnJ], d[kLnJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
      const int c4 = lane + j * 32;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < n4) {
        x[j] = __ldg(reinterpret_cast<const float4 *>(u + row * N) + c4);
        d[j] = __ldg(reinterpret_cast<const float4 *>(dy + row * N) + c4);
        sum += (x[j].x + x[j].y) + (x[j].z + x[j].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
    const float mean = sum * invN;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < n4) {
        x[j].x -= mean; x[j].y -= mean; x[j].z -= mean; x[j].w -= mean;
        sq = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, fmaf(x[j].w, x[j].w, sq))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFullMask, sq, o);
    const float rstd = 1.0f / sqrtf(sq * invN + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < n4) {
        x[j].x *= rstd; x[j].y *= rstd; x[j].z *= rstd; x[j].w *= rstd;  // xhat
        ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
        ag[j].x = fmaf(d[j].x, x[j].x, ag[j].x); ag[j].y = fmaf(d[j].y, x[j].y, ag[j].y);
        ag[j].z = fmaf(d[j].z, x[j].z, ag[j].z); ag[j].w = fmaf(d[j].w, x[j].w, ag[j].w);
        d[j].x *= g4[j].x; d[j].y *= g4[j].y; d[j].z *= g4[j].z; d[j].w *= g4[j].w;  // dy * gamma
        c1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
        c2 = fmaf(d[j].x, x[j].x, fmaf(d[j].y, x[j].y, fmaf(d[j].z, x[j].z, fmaf(d[j].w, x[j].w, c2))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c1 += __shfl_xor_sync(kFullMask, c1, o);
      c2 += __shfl_xor_sync(kFullMask, c2, o);
    }
    c1 *= invN;
    c2 *= invN;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < n4) {
        float4 o;
        o.x = rstd * (d[j].x - c1 - x[j].x * c2);
        o.y = rstd * (d[j].y - c1 - x[j].y * c2);
        o.z = rstd * (d[j].z - c1 - x[j].z * c2);
        o.w = rstd * (d[j].w - c1 - x[j].w * c2);
        reinterpret_cast<float4 *>(du + row * N)[c4] = o;
        if (dproj) {
          // the forward GEMM epilogue hashed (row * 3 + problem 0, column): linear.cu, kMaxProbs = 3
          const uint32_t ra = (uint32_t)row * 3u, gc = (uint32_t)(c4 * 4);
          o.x = dropout_keep(drop_seed, ra, gc + 0u, drop_thresh) ? o.x * drop_scale : 0.f;
          o.y = dropout_keep(drop_seed, ra, gc + 1u, drop_thresh) ? o.y * drop_scale : 0.f;
          o.z = dropout_keep(drop_seed, ra, gc + 2u, drop_thresh) ? o.z * drop_scale : 0.f;
          o.w = dropout_keep(drop_seed, ra, gc + 3u, drop_thresh) ? o.w * drop_scale : 0.f;
          reinterpret_cast<float4 *>(dproj + row * N)[c4] = o;
        }
      }
    }
  }
  // per-column partial sums: warps -> shared -> global
#pragma unroll
  for (int j = 0; j < kLnJ; ++j) {
    const int c = (lane + j * 32) * 4;
    if (c < N) {
      atomicAdd(&s_acc[0][c + 0], ag[j].x); atomicAdd(&s_acc[0][c + 1], ag[j].y);
      atomicAdd(&s_acc[0][c + 2], ag[j].z); atomicAdd(&s_acc[0][c + 3], ag[j].w);
      atomicAdd(&s_acc[1][c + 0], ab[j].x); atomicAdd(&s_acc[1][c + 1], ab[j].y);
      atomicAdd(&s_acc[1][c + 2], ab[j].z); atomicAdd(&s_acc[1][c + 3], ab[j].w);
    }
  }
  __syncthreads();
  for (int c = tid; c < N; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, s_acc[0][c]);
    if (dbeta) atomicAdd(dbeta + c, s_acc[1][c]);
  }
}

__global__ void relu_backward_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ y, float scale,
                                     long long n4, float4 *__restrict__ out) {
  pdl_launch_dependents();  // a PDL-launched successor may be scheduled now (it waits for this grid's completion itself)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 d = __ldg(dy + i), a = __ldg(y + i);
    out[i] = make_float4(a.x > 0.f ? d.x * scale : 0.f, a.y > 0.f ? d.y * scale : 0.f, a.z > 0.f ? d.z * scale : 0.f,
                         a.w > 0.f ? d.w * scale : 0.f);
  }
}


// ---------------------------------------------------------------------------------------------------------
// weight gradient for tiny K (the 3- or 6-channel input of PositionEmbeddingLearned's first 1x1 conv,
// encoder_decoder_layers.py:24-28): dW[N x K] += dY^T X, db[N] += column sums of dY, K <= 8, plain fp32 FMAs.
// Block = 32 output rows n x 8 k lanes over one 256-row chunk: dY reads are coalesced over n, X reads broadcast.
// ---------------------------------------------------------------------------------------------------------
constexpr int kWsRows = 64;  // rows per block: the loop is a chain of dependent global loads, so many short blocks
                             // (2048 rows x 288 outputs: 288 blocks) instead of few long ones (46 us -> a few us)
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const float *__restrict__ dy, int ldy, const float *__restrict__ x, int ldx, long long rows, int N,
                   int K, float *__restrict__ dw, int ldw, float *__restrict__ db) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int k = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * kWsRows;
  const long long r1 = (r0 + kWsRows < rows) ? r0 + kWsRows : rows;
  if (n >= N) return;
  float acc = 0.f, sum = 0.f;
  const bool kk = k < K;
#pragma unroll 8
  for (long long r = r0; r < r1; ++r) {
    const float g = __ldg(dy + r * ldy + n);
    if (kk) acc = fmaf(g, __ldg(x + r * ldx + k), acc);
    sum += g;
  }
  if (kk) atomicAdd(dw + (long long)n * ldw + k, acc);
  if (k == 0 && db) atomicAdd(db + n, sum);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace eda

extern "C" {

int eda_wgrad(const eda_wgrad_problem *probs, int nprobs, int N, int K, void *stream) {
  using namespace eda;
  if (!probs || nprobs < 1 || nprobs > kWgMaxProbs || N < 1 || K < 1) return EDA_ERR_INVALID_ARGUMENT;
  if ((N & 3) || (K & 3)) return EDA_ERR_UNSUPPORTED;
  WgParams p = {};
  long long max_rows = 0;
  for (int i = 0; i < nprobs; ++i) {
    const eda_wgrad_problem &q = probs[i];
    if (q.rows < 0) return EDA_ERR_INVALID_ARGUMENT;
    if (q.rows > 0 && (!q.dy || !q.x || !q.dw)) return EDA_ERR_INVALID_ARGUMENT;
    if (q.ldy < N || q.ldx < K || q.ldw < K || (q.ldy & 3) || (q.ldx & 3) || !aligned16(q.dy) || !aligned16(q.x))
      return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].dy = q.dy; p.pr[i].x = q.x; p.pr[i].dw = q.dw; p.pr[i].db = q.db; p.pr[i].rows = q.rows;
    p.pr[i].x_scale = q.x_scale; p.pr[i].x_shift = q.x_shift;
    if ((q.x_scale == nullptr) != (q.x_shift == nullptr)) return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].ldy = q.ldy; p.pr[i].ldx = q.ldx; p.pr[i].ldw = q.ldw;
    if (q.rows > max_rows) max_rows = q.rows;
  }
  if (max_rows == 0) return EDA_OK;
  // large row counts without an input prologue: the tcgen05 kernel (wgrad_tc.cu), operands by TMA as they lie in memory
  if (wgrad_tc_eligible(probs, nprobs, N, K)) {
    const int rc = wgrad_tc_launch(probs, nprobs, N, K, as_stream(stream));
    if (rc != kWgradTcDeclined) return rc;
  }
  p.nprobs = nprobs; p.N = N; p.K = K;
  const int nt = (N + kWgTile - 1) / kWgTile, kt = (K + kWgTile - 1) / kWgTile;
  const int sms = sm_count();
  // enough row splits to fill the chip a few times over, but at least 128 rows per split
  long long splits = (4LL * sms + (long long)nt * kt * nprobs - 1) / ((long long)nt * kt * nprobs);
  const long long max_splits = (max_rows + 127) / 128;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits * nprobs > 65535) splits = 65535 / nprobs;
  p.splits = (int)splits;
  dim3 grid((unsigned)nt, (unsigned)kt, (unsigned)(p.splits * nprobs));
  wgrad_kernel<<<grid, kWgThreads, 0, as_stream(stream)>>>(p);
  return check_launch("wgrad_kernel");
}

int eda_layernorm_backward(const float *dy, const float *u, const float *gamma, float eps, long long rows, int N,
                           float *du, float *dproj, float *dgamma, float *dbeta, float dropout_p,
                           unsigned int dropout_seed, const unsigned int *dropout_epoch, void *stream) {
  using namespace eda;
  if (rows < 0 || N < 4 || (N & 3) || N > kLnMaxN) return rows < 0 ? EDA_ERR_INVALID_ARGUMENT : EDA_ERR_UNSUPPORTED;
  if (rows == 0) return EDA_OK;
  if (!dy || !u || !du || dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  if (!aligned16(dy) || !aligned16(u) || !aligned16(du) || (dproj && !aligned16(dproj)) || (gamma && !aligned16(gamma)))
    return EDA_ERR_INVALID_ARGUMENT;
  if (rows * 3 > 0xffffffffLL) return EDA_ERR_UNSUPPORTED;
  // one row per warp until the chip is full (one block per SM), more rows per warp beyond that: the per-row chain (two
  // loads, two warp reductions, a store) is pure latency, and the 2 N atomics per block at the end are cheap next to it
  // (scripts/ln_bwd_time.py: 640 / 2048 / 8192 rows 11.4 / 11.6 / 15.4 us with >= 8 rows per warp, 5.1 / 5.6 / 10.9 us so)
  long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int sms = sm_count();
  if (blocks > sms) blocks = sms;
  if (blocks < 1) blocks = 1;
  const uint32_t thresh = dropout_thresh(dropout_p);
  EDA_CUDA_TRY(launch_pdl(layernorm_backward_kernel, dim3((unsigned)blocks), dim3(kLnWarps * 32), 0, as_stream(stream), dy, u,
                          gamma, eps, rows, N, du, thresh ? dproj : nullptr, dgamma, dbeta, thresh, dropout_seed,
                          1.0f / (1.0f - dropout_p), reinterpret_cast<const uint32_t *>(dropout_epoch)),
               "layernorm_backward_kernel launch");
  return check_launch("layernorm_backward_kernel");
}

int eda_relu_backward(const float *dy, const float *y, float scale, long long n, float *out, void *stream) {
  using namespace eda;
  if (n < 0 || (n & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (n == 0) return EDA_OK;
  if (!dy || !y || !out || !aligned16(dy) || !aligned16(y) || !aligned16(out)) return EDA_ERR_INVALID_ARGUMENT;
  const long long n4 = n >> 2;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  relu_backward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(dy), reinterpret_cast<const float4 *>(y), scale, n4, reinterpret_cast<float4 *>(out));
  return check_launch("relu_backward_kernel");
}


int eda_wgrad_small(const float *dy, int ldy, const float *x, int ldx, long long rows, int N, int K, float *dw, int ldw,
                    float *db, void *stream) {
  using namespace eda;
  if (rows < 0 || N < 1 || K < 1 || ldy < N || ldx < K || ldw < K) return EDA_ERR_INVALID_ARGUMENT;
  if (K > 8) return EDA_ERR_UNSUPPORTED;
  if (rows == 0) return EDA_OK;
  if (!dy || !x || !dw) return EDA_ERR_INVALID_ARGUMENT;
  const long long chunks = (rows + kWsRows - 1) / kWsRows;
  if (chunks > 65535) return EDA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)chunks);
  wgrad_small_kernel<<<grid, 256, 0, as_stream(stream)>>>(dy, ldy, x, ldx, rows, N, K, dw, ldw, db);
  return check_launch("wgrad_small_kernel");
}

}  // extern "C"
