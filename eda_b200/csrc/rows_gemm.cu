// Row-streaming GEMM for the set-abstraction backward pass (sm_100a):
//
//     Y[R x N] = f(X)[R x K] W'[N x K]^T,   f(x) = x   or   relu(x * in_scale[k] + in_shift[k])
//
// R is 10^5 - 10^6 grouped points, K and N are the MLP widths (16 ... 272): every operand row is touched once, so
// the kernel is bound by HBM (4 (K + N) bytes per row), not by the tensor pipe, and what matters is that rows stream
// at full bandwidth: persistent CTAs (one per SM), the weight slice (<= 128 output channels) converted to tf32 once
// into shared memory, X tiles of 128 rows x 32 channels in a 3-stage cp.async ring that runs across tile
// boundaries, warp-level mma.sync.m16n8k8 tf32 with fp32 accumulators (warp = 16 rows x up to 128 columns), results
// written as full 32-byte sectors.  The optional prologue f applies the previous layer's folded BatchNorm + ReLU while
// the A fragments are formed, so the post-activation tensor never exists in HBM; W' is addressed through two strides,
// so W^T (activation gradients, dX = dY W) needs no transposed copy.
// Measured (profiles/r1_bwd_kernel_metrics.json): 120-130 us per launch, 1.9-2.4 TB/s, legacy tensor pipe 22-29 %, issue
// slots 40 % active — latency-bound with one 8-warp CTA per SM, not yet at the HBM roofline it is designed for.
#include "common.cuh"
#include "rows_gemm_tc.h"

namespace eda {
namespace {

constexpr int kRgRows = 128, kRgThreads = 256, kRgKc = 32, kRgPx = kRgKc + 4, kRgStages = 3, kRgNc = 128;
constexpr int kRgMaxK = 288;

struct RowsGemmParams {
  const float *x, *in_scale, *in_shift, *w;
  float *y;
  double *stats;  // optional [sum_r y[r][n] (N), sum_r y[r][n]^2 (N)], accumulated (fp64 atomics): the BatchNorm batch
                  // statistics of the layer this GEMM evaluates, taken in the epilogue instead of by a second pass over y
  long long rows, w_sn, w_sk;
  int ldx, ldy, K, N, ntiles;
};

__device__ __forceinline__ uint32_t rg_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void rg_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void rg_cp16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}

template <bool kPrologue, bool kStats>
__global__ void __launch_bounds__(kRgThreads, 2)
rows_gemm_kernel(const RowsGemmParams p) {
  extern __shared__ __align__(16) unsigned char rg_smem[];
  const int K = p.K, PW = K + 4;  // (g * PW + t) % 32 = 4 g + t: conflict-free B fragments (K % 8 == 0)
  const int n0 = blockIdx.y * kRgNc;
  const int Nc = min(kRgNc, p.N - n0);         // multiple of 8
  uint32_t *sW = reinterpret_cast<uint32_t *>(rg_smem);                    // [Nc][PW] tf32
  float *sX = reinterpret_cast<float *>(sW + (size_t)kRgNc * PW);          // [stages][128][36]
  float *sSc = sX + kRgStages * kRgRows * kRgPx;                            // [K] in_scale, [K] in_shift
  float *sSh = sSc + kRgMaxK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  const int nkc = (K + kRgKc - 1) / kRgKc;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const long long total = (long long)my_tiles * nkc;  // (tile, k-chunk) items of this CTA, in order

  auto issue = [&](long long item) {
    if (item < total) {
      const int ti = (int)(item / nkc), kc = (int)(item - (long long)ti * nkc);
      const long long row0 = ((long long)blockIdx.x + (long long)ti * gridDim.x) * kRgRows;
      float *dst = sX + (size_t)(item % kRgStages) * kRgRows * kRgPx;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int id = i * kRgThreads + tid;
        const int r = id >> 3, c4 = (id & 7) << 2;
        const long long row = row0 + r;
        const int k = kc * kRgKc + c4;
        const bool in = row < p.rows && k < K;
        rg_cp16(dst + r * kRgPx + c4, in ? p.x + row * p.ldx + k : p.x, in ? 16u : 0u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  issue(1);

  // weights of this column slice -> shared memory (tf32), prologue vectors
  for (int i = tid; i < Nc * K; i += kRgThreads) {
    const int n = i / K, k = i - n * K;
    sW[n * PW + k] = rg_tf32(__ldg(p.w + (n0 + n) * p.w_sn + k * p.w_sk));
  }
  if (kPrologue)
    for (int i = tid; i < K; i += kRgThreads) {
      sSc[i] = __ldg(p.in_scale + i);
      sSh[i] = __ldg(p.in_shift + i);
    }

  constexpr int kMaxNT = kRgNc / 8;
  const int nnt = Nc >> 3;
  float acc[kMaxNT][4];
  // column statistics (kStats): this thread's running partial sums over all the rows it produces, for its columns
  // j * 8 + 2 t and + 1 — fixed order, fp32; reduced over the CTA at the end and added to p.stats in fp64
  float csum[kStats ? kMaxNT : 1][2], csq[kStats ? kMaxNT : 1][2];
  if (kStats) {
#pragma unroll
    for (int j = 0; j < kMaxNT; ++j) csum[j][0] = csum[j][1] = csq[j][0] = csq[j][1] = 0.f;
  }
  const int r0 = warp * 16;
  for (long long item = 0; item < total; ++item) {
    const int ti = (int)(item / nkc), kc = (int)(item - (long long)ti * nkc);
    issue(item + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();  // chunk `item` has landed for every thread (first iteration: also the weights)
    if (kc == 0) {
#pragma unroll
      for (int j = 0; j < kMaxNT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    }
    const float *xs = sX + (size_t)(item % kRgStages) * kRgRows * kRgPx;
    const int kcnt = min(kRgKc, K - kc * kRgKc);
    for (int ks = 0; ks < kcnt / 8; ++ks) {
      const int kl = ks * 8 + t, kg = kc * kRgKc + kl;
      float a0 = xs[(r0 + g) * kRgPx + kl], a1 = xs[(r0 + g + 8) * kRgPx + kl];
      float a2 = xs[(r0 + g) * kRgPx + kl + 4], a3 = xs[(r0 + g + 8) * kRgPx + kl + 4];
      if (kPrologue) {
        const float s0 = sSc[kg], h0 = sSh[kg], s1 = sSc[kg + 4], h1 = sSh[kg + 4];
        a0 = fmaxf(fmaf(a0, s0, h0), 0.f); a1 = fmaxf(fmaf(a1, s0, h0), 0.f);
        a2 = fmaxf(fmaf(a2, s1, h1), 0.f); a3 = fmaxf(fmaf(a3, s1, h1), 0.f);
      }
      const uint32_t a[4] = {rg_tf32(a0), rg_tf32(a1), rg_tf32(a2), rg_tf32(a3)};
      const uint32_t *wr = sW + g * PW + kg;
#pragma unroll
      for (int j = 0; j < kMaxNT; ++j)
        if (j < nnt) rg_mma(acc[j], a, wr[j * 8 * PW], wr[j * 8 * PW + 4]);
    }
    if (kc == nkc - 1) {
      const long long row0 = ((long long)blockIdx.x + (long long)ti * gridDim.x) * kRgRows;
      const long long ra = row0 + r0 + g, rb = ra + 8;
      const float fa = ra < p.rows ? 1.f : 0.f, fb = rb < p.rows ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < kMaxNT; ++j) {
        if (j < nnt) {
          const int n = n0 + j * 8 + 2 * t;
          if (ra < p.rows) *reinterpret_cast<float2 *>(p.y + ra * p.ldy + n) = make_float2(acc[j][0], acc[j][1]);
          if (rb < p.rows) *reinterpret_cast<float2 *>(p.y + rb * p.ldy + n) = make_float2(acc[j][2], acc[j][3]);
          if (kStats) {  // rows beyond p.rows do not count (zero-filled on the way in, but f(0) = relu(shift) != 0)
            const float v0 = acc[j][0] * fa, v1 = acc[j][1] * fa;
            const float v2 = acc[j][2] * fb, v3 = acc[j][3] * fb;
            csum[j][0] += v0 + v2;
            csum[j][1] += v1 + v3;
            csq[j][0] = fmaf(v0, v0, fmaf(v2, v2, csq[j][0]));
            csq[j][1] = fmaf(v1, v1, fmaf(v3, v3, csq[j][1]));
          }
        }
      }
    }
    __syncthreads();  // the stage read here is refilled by the next iteration's prefetch
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (kStats) {
    // lanes with the same t hold the same columns: butterfly over g (lane bits 2..4), then the 8 warps through shared
    // memory (the X ring is idle now), then ONE fp64 atomic per column and CTA
    __syncthreads();
    float *red = sX;  // [8 warps][2 * kRgNc]
#pragma unroll
    for (int j = 0; j < kMaxNT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = csum[j][e], b = csq[j][e];
#pragma unroll
        for (int d = 4; d < 32; d <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, d);
          b += __shfl_xor_sync(0xffffffffu, b, d);
        }
        if (g == 0 && j < nnt) {
          red[warp * 2 * kRgNc + j * 8 + 2 * t + e] = a;
          red[warp * 2 * kRgNc + kRgNc + j * 8 + 2 * t + e] = b;
        }
      }
    }
    __syncthreads();
    for (int c = tid; c < 2 * Nc; c += kRgThreads) {
      const int col = c < Nc ? c : c - Nc, half = c < Nc ? 0 : 1;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kRgThreads / 32; ++w) v += (double)red[w * 2 * kRgNc + half * kRgNc + col];
      atomicAdd(p.stats + (size_t)half * p.N + n0 + col, v);
    }
  }
}

}  // namespace
}  // namespace eda

extern "C" int eda_rows_gemm_stats(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                                   long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y,
                                   int ldy, double *stats, void *stream);

extern "C" int eda_rows_gemm(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                             long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y, int ldy,
                             void *stream) {
  return eda_rows_gemm_stats(x, ldx, in_scale, in_shift, w, w_stride_n, w_stride_k, rows, K, N, y, ldy, nullptr, stream);
}

extern "C" int eda_rows_gemm_stats(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                                   long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y,
                                   int ldy, double *stats, void *stream) {
  using namespace eda;
  if (rows < 0 || K < 8 || N < 8) return EDA_ERR_INVALID_ARGUMENT;
  if ((K & 7) || (N & 7) || K > kRgMaxK) return EDA_ERR_UNSUPPORTED;
  if (rows == 0) return EDA_OK;
  if (!x || !w || !y || ldx < K || ldy < N || (ldx & 3) || (ldy & 1) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 7) || ((in_scale == nullptr) != (in_shift == nullptr)))
    return EDA_ERR_INVALID_ARGUMENT;
  const long long ntiles = (rows + kRgRows - 1) / kRgRows;
  if (ntiles > 0x7fffffffLL) return EDA_ERR_UNSUPPORTED;
  // the big launches (K <= 160, >= 16k rows): persistent tcgen05 kernel, operands in and out by TMA (rows_gemm_tc.cu)
  if (rows_gemm_tc_eligible(x, ldx, rows, K, N, y, ldy)) {
    const int rc = rows_gemm_tc_launch(x, ldx, in_scale, in_shift, w, w_stride_n, w_stride_k, rows, K, N, y, ldy, stats,
                                       as_stream(stream));
    if (rc != kRowsGemmTcDeclined) return rc;
  }
  RowsGemmParams p = {};
  p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.y = y; p.rows = rows; p.w_sn = w_stride_n;
  p.w_sk = w_stride_k; p.ldx = ldx; p.ldy = ldy; p.K = K; p.N = N; p.ntiles = (int)ntiles; p.stats = stats;
  const size_t smem = (size_t)kRgNc * (K + 4) * 4 + (size_t)kRgStages * kRgRows * kRgPx * 4 + 2 * kRgMaxK * 4;
  static SmemAttr attr[4];
  const int which = (in_scale ? 1 : 0) | (stats ? 2 : 0);
  switch (which) {
    case 0: EDA_CUDA_TRY(attr[0].ensure(rows_gemm_kernel<false, false>, smem), "rows_gemm smem attr"); break;
    case 1: EDA_CUDA_TRY(attr[1].ensure(rows_gemm_kernel<true, false>, smem), "rows_gemm smem attr"); break;
    case 2: EDA_CUDA_TRY(attr[2].ensure(rows_gemm_kernel<false, true>, smem), "rows_gemm smem attr"); break;
    default: EDA_CUDA_TRY(attr[3].ensure(rows_gemm_kernel<true, true>, smem), "rows_gemm smem attr"); break;
  }
  const int sms = sm_count();
  const int ny = (N + kRgNc - 1) / kRgNc;
  // two resident CTAs per SM when the weight slice leaves room (K <= 64: 92 KB each): the kernel is latency-bound with
  // one 8-warp CTA (issue slots 40 % active), the second CTA fills its stalls.  EDA_ROWS_GEMM_CTAS=1 keeps one.
  static const int max_ctas = [] { const char *e = getenv("EDA_ROWS_GEMM_CTAS"); return e ? atoi(e) : 2; }();
  const int per_sm = (2 * (smem + 1024) <= 227 * 1024 && max_ctas >= 2) ? 2 : 1;
  long long gx = (long long)per_sm * sms / ny;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dim3 grid((unsigned)gx, (unsigned)ny);
  switch (which) {
    case 0: rows_gemm_kernel<false, false><<<grid, kRgThreads, smem, as_stream(stream)>>>(p); break;
    case 1: rows_gemm_kernel<true, false><<<grid, kRgThreads, smem, as_stream(stream)>>>(p); break;
    case 2: rows_gemm_kernel<false, true><<<grid, kRgThreads, smem, as_stream(stream)>>>(p); break;
    default: rows_gemm_kernel<true, true><<<grid, kRgThreads, smem, as_stream(stream)>>>(p); break;
  }
  return check_launch("rows_gemm_kernel");
}
