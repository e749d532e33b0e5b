// Row-tiled linear layer with fused epilogues for the cross-modal attention layers (sm_100a).
//
//     Y[R x N] = act( (X [+ P])[R x K] * W[N x K]^T + bias )                       (plain)
//     Y        = LayerNorm( residual + (X [+ P]) W^T + bias ) * gamma + beta        (ln epilogue)
//
// Replaces the nn.Linear / F.linear calls inside nn.MultiheadAttention's math path (in-projection
// of q, k, v and the out-projection, torch/nn/functional.py:6607-6665 as used by
// models/encoder_decoder_layers.py:47-71,133,298-319), the residual + nn.LayerNorm that follows every
// attention block (encoder_decoder_layers.py:94-96,106-107,118-122,371,381,392,402), the two FFN
// linears (:53-59, :322-328) and the 1x1 Conv1d's of PositionEmbeddingLearned (:24-28).
// The reference issues 3-6 library kernels per such block; here the bias, "+ pos" on the input,
// ReLU, residual add and LayerNorm all ride in the GEMM's prologue / epilogue.
//
// Up to three independent problems (same K, N, epilogue) share one launch: the q / k / v
// in-projections of one attention block are one grid.
//
// CTA = one 128-row tile.  K is streamed in 32-wide blocks through a 3-stage shared-memory ring:
//   A block  : staged by the 128 compute threads (thread = row): 16-byte global loads, "+ pos",
//              cvt.rna.tf32, float4 stores in the K-major core-matrix layout (conflict-free)
//   W block  : pre-packed (eda_linear_pack), one 1-D TMA bulk copy by the producer warp
//   MMA      : tcgen05.mma kind::tf32, M = 128, N <= 256 per instruction (N = 288 -> 2 x 144),
//              fp32 accumulator 128 x N in tensor memory
// Epilogue: thread = row = TMEM lane; the LayerNorm variant makes three passes over its TMEM row
// (sum -> centred sum of squares -> normalise), so no shared memory or shuffles are involved.
#include "umma.cuh"

namespace eda {
namespace {

constexpr int kRows = 128;
constexpr int kThreads = 160;  // warps 0-3 compute (thread = row), warp 4 = weight producer
constexpr int kStages = 3;
constexpr int kKBlock = 32;
constexpr int kABytes = kRows * kKBlock * 4;  // 16 KB
constexpr int kMaxN = 320;
constexpr int kMaxProbs = 3;

struct LinProblem {
  const float *x, *pos, *w, *bias, *residual;
  float *y;
  int rows, tile0;
};

struct LinParams {
  LinProblem pr[kMaxProbs];
  int nprobs, K, Kpad, N, relu, ln;
  const float *gamma, *beta;
  float eps;
  uint32_t tmem_cols, stage_bytes;
};

__device__ __forceinline__ void named_bar_sync_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const LinParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_w[kStages], empty[kStages], mma_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[kMaxN], s_gamma[kMaxN], s_beta[kMaxN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pi = 0;
  while (pi + 1 < p.nprobs && (int)blockIdx.x >= p.pr[pi + 1].tile0) ++pi;
  const LinProblem &pr = p.pr[pi];
  const int tile = (int)blockIdx.x - pr.tile0;
  const int N = p.N, K = p.K;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&mma_done, 1);
    mbar_fence_init_cluster();
  }
  for (int i = tid; i < N; i += kThreads) {
    s_bias[i] = pr.bias ? __ldg(pr.bias + i) : 0.f;
    s_gamma[i] = (p.ln && p.gamma) ? __ldg(p.gamma + i) : 1.f;
    s_beta[i] = (p.ln && p.beta) ? __ldg(p.beta + i) : 0.f;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  const int nkb = (p.Kpad + kKBlock - 1) / kKBlock;

  if (warp == 4) {
    // ---------------- weight producer ----------------------------------------------------------
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % kStages;
        const int kcnt = min(kKBlock, p.Kpad - kb * kKBlock);
        const uint32_t bytes = (uint32_t)kcnt * (uint32_t)N * 4u;
        mbar_wait(&empty[slot], ((kb / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_w[slot], bytes);
        bulk_g2s(smem_raw + (size_t)slot * p.stage_bytes + kABytes, pr.w + (size_t)kb * kKBlock * N, bytes,
                 &full_w[slot]);
      }
    }
  } else {
    // ---------------- compute warps: thread = row ------------------------------------------------
    const long long row = (long long)tile * kRows + tid;
    const bool valid = row < pr.rows;
    const float *xrow = pr.x + row * K;
    const float *prow = pr.pos ? pr.pos + row * K : nullptr;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(pr.x) & 15) == 0 &&
                     (!pr.pos || (reinterpret_cast<uintptr_t>(pr.pos) & 15) == 0);
    // N split into MMA-sized pieces (multiples of 16, <= 256)
    const int n_a = N <= 256 ? N : ((N / 2 + 15) / 16) * 16;
    const int n_b = N - n_a;

    for (int kb = 0; kb < nkb; ++kb) {
      const int slot = kb % kStages;
      const int kcnt = min(kKBlock, p.Kpad - kb * kKBlock);
      const int nch = kcnt >> 2;
      float4 v[kKBlock / 4];
#pragma unroll
      for (int c = 0; c < kKBlock / 4; ++c) {
        v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = kb * kKBlock + c * 4;
        if (c < nch && valid && k < K) {
          if (vec) {
            v[c] = __ldg(reinterpret_cast<const float4 *>(xrow + k));
            if (prow) {
              const float4 q = __ldg(reinterpret_cast<const float4 *>(prow + k));
              v[c].x += q.x; v[c].y += q.y; v[c].z += q.z; v[c].w += q.w;
            }
          } else {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[e] = 0.f;
              if (k + e < K) f[e] = __ldg(xrow + k + e) + (prow ? __ldg(prow + k + e) : 0.f);
            }
            v[c] = make_float4(f[0], f[1], f[2], f[3]);
          }
        }
      }
      mbar_wait(&empty[slot], ((kb / kStages) & 1) ^ 1);  // the MMAs that read this slot have finished
      float4 *sA = reinterpret_cast<float4 *>(smem_raw + (size_t)slot * p.stage_bytes);
#pragma unroll
      for (int c = 0; c < kKBlock / 4; ++c)
        if (c < nch)
          sA[c * kRows + tid] = make_float4(to_tf32(v[c].x), to_tf32(v[c].y), to_tf32(v[c].z), to_tf32(v[c].w));
      umma::fence_proxy_async_smem();
      umma::fence_before_thread_sync();
      named_bar_sync_compute();
      if (tid == 0) {
        mbar_wait(&full_w[slot], (kb / kStages) & 1);
        umma::fence_after_thread_sync();
        const uint32_t abase = smem_u32(sA);
        const uint32_t wbase = abase + kABytes;
        const uint32_t lbo_w = (uint32_t)N * 16u;
        for (int ks = 0; ks < kcnt / 8; ++ks) {
          const uint64_t adesc = umma::smem_desc_kmajor_noswizzle(abase + (uint32_t)ks * 2u * kRows * 16u, kRows * 16u, 128u);
          const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
          const uint64_t b0 = umma::smem_desc_kmajor_noswizzle(wbase + (uint32_t)ks * 2u * lbo_w, lbo_w, 128u);
          umma::mma_tf32_ss(tbase, adesc, b0, umma::idesc_tf32(kRows, n_a), acc);
          if (n_b > 0) {
            const uint64_t b1 =
                umma::smem_desc_kmajor_noswizzle(wbase + (uint32_t)ks * 2u * lbo_w + (uint32_t)n_a * 16u, lbo_w, 128u);
            umma::mma_tf32_ss(tbase + (uint32_t)n_a, adesc, b1, umma::idesc_tf32(kRows, n_b), acc);
          }
        }
        umma::mma_commit(&empty[slot]);
        if (kb == nkb - 1) umma::mma_commit(&mma_done);
      }
    }
    mbar_wait(&mma_done, 0);
    umma::fence_after_thread_sync();
    __syncwarp();

    // ---------------- epilogue -------------------------------------------------------------------
    const uint32_t trow = umma::tmem_addr(tbase, (uint32_t)(warp * 32), 0);
    float *yrow = pr.y + row * N;
    if (!p.ln) {
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(trow + (uint32_t)c0, u);
        umma::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 o;
            o.x = __uint_as_float(u[q4 * 4 + 0]) + s_bias[c0 + q4 * 4 + 0];
            o.y = __uint_as_float(u[q4 * 4 + 1]) + s_bias[c0 + q4 * 4 + 1];
            o.z = __uint_as_float(u[q4 * 4 + 2]) + s_bias[c0 + q4 * 4 + 2];
            o.w = __uint_as_float(u[q4 * 4 + 3]) + s_bias[c0 + q4 * 4 + 3];
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(yrow + c0 + q4 * 4) = o;
          }
        }
      }
    } else {
      const float *rrow = pr.residual ? pr.residual + row * N : nullptr;
      float sum = 0.f;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(trow + (uint32_t)c0, u);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && rrow) r4 = __ldg(reinterpret_cast<const float4 *>(rrow + c0 + q4 * 4));
          const float a0 = __uint_as_float(u[q4 * 4 + 0]) + s_bias[c0 + q4 * 4 + 0] + r4.x;
          const float a1 = __uint_as_float(u[q4 * 4 + 1]) + s_bias[c0 + q4 * 4 + 1] + r4.y;
          const float a2 = __uint_as_float(u[q4 * 4 + 2]) + s_bias[c0 + q4 * 4 + 2] + r4.z;
          const float a3 = __uint_as_float(u[q4 * 4 + 3]) + s_bias[c0 + q4 * 4 + 3] + r4.w;
          sum += (a0 + a1) + (a2 + a3);
          u[q4 * 4 + 0] = __float_as_uint(a0); u[q4 * 4 + 1] = __float_as_uint(a1);
          u[q4 * 4 + 2] = __float_as_uint(a2); u[q4 * 4 + 3] = __float_as_uint(a3);
        }
        umma::tmem_st16(trow + (uint32_t)c0, u);
      }
      umma::tmem_st_wait();
      const float mean = sum / (float)N;
      float ss = 0.f;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(trow + (uint32_t)c0, u);
        umma::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float d = __uint_as_float(u[e]) - mean;
          ss = fmaf(d, d, ss);
        }
      }
      const float rstd = 1.0f / sqrtf(ss / (float)N + p.eps);
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(trow + (uint32_t)c0, u);
        umma::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 o;
            o.x = (__uint_as_float(u[q4 * 4 + 0]) - mean) * rstd * s_gamma[c0 + q4 * 4 + 0] + s_beta[c0 + q4 * 4 + 0];
            o.y = (__uint_as_float(u[q4 * 4 + 1]) - mean) * rstd * s_gamma[c0 + q4 * 4 + 1] + s_beta[c0 + q4 * 4 + 1];
            o.z = (__uint_as_float(u[q4 * 4 + 2]) - mean) * rstd * s_gamma[c0 + q4 * 4 + 2] + s_beta[c0 + q4 * 4 + 2];
            o.w = (__uint_as_float(u[q4 * 4 + 3]) - mean) * rstd * s_gamma[c0 + q4 * 4 + 3] + s_beta[c0 + q4 * 4 + 3];
            *reinterpret_cast<float4 *>(yrow + c0 + q4 * 4) = o;
          }
        }
      }
    }
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, p.tmem_cols);
}

// W (N, K) row-major [* scale[n]] -> blocks [kb] of float4 [kcnt/4][N] (K-major core-matrix layout,
// chunk-major), rounded to tf32 once.  K is zero-padded to a multiple of 8.
__global__ void pack_linear_kernel(const float *__restrict__ W, const float *__restrict__ scale, int N, int K, int Kpad,
                                   float *__restrict__ dst) {
  const int total = N * Kpad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kb = e / (kKBlock * N);
    const int rem = e - kb * kKBlock * N;
    const int c = rem / (4 * N);
    const int n = (rem >> 2) % N;
    const int k = kb * kKBlock + c * 4 + (rem & 3);
    float w = 0.f;
    if (k < K) {
      w = W[(size_t)n * K + k];
      if (scale) w *= scale[n];
    }
    dst[e] = to_tf32(w);
  }
}

inline int kpad_of(int K) { return (K + 7) & ~7; }
inline bool lin_supported(int N, int K) { return N >= 16 && N <= kMaxN && (N & 15) == 0 && K >= 1 && K <= 4096; }

}  // namespace
}  // namespace eda

extern "C" {

size_t eda_linear_packed_floats(int N, int K) {
  if (!eda::lin_supported(N, K)) return 0;
  return (size_t)N * eda::kpad_of(K);
}

int eda_linear_pack(const float *W, const float *scale, int N, int K, float *packed, void *stream) {
  using namespace eda;
  if (!lin_supported(N, K)) return EDA_ERR_UNSUPPORTED;
  if (!W || !packed) return EDA_ERR_INVALID_ARGUMENT;
  const int total = N * kpad_of(K);
  pack_linear_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(W, scale, N, K, kpad_of(K), packed);
  return check_launch("pack_linear_kernel");
}

int eda_linear_forward(const eda_linear_problem *probs, int nprobs, int K, int N, int relu, const float *ln_gamma,
                       const float *ln_beta, float ln_eps, int layer_norm, void *stream) {
  using namespace eda;
  if (!probs || nprobs < 1 || nprobs > kMaxProbs) return EDA_ERR_INVALID_ARGUMENT;
  if (!lin_supported(N, K)) return EDA_ERR_UNSUPPORTED;
  LinParams p = {};
  int tiles = 0;
  for (int i = 0; i < nprobs; ++i) {
    const eda_linear_problem &q = probs[i];
    if (q.rows < 0) return EDA_ERR_INVALID_ARGUMENT;
    if (q.rows > 0 && (!q.x || !q.w_packed || !q.y)) return EDA_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(q.w_packed) & 15) || (reinterpret_cast<uintptr_t>(q.y) & 15) ||
        (q.residual && (reinterpret_cast<uintptr_t>(q.residual) & 15)))
      return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].x = q.x; p.pr[i].pos = q.pos; p.pr[i].w = q.w_packed; p.pr[i].bias = q.bias;
    p.pr[i].residual = q.residual; p.pr[i].y = q.y; p.pr[i].rows = q.rows; p.pr[i].tile0 = tiles;
    tiles += (q.rows + kRows - 1) / kRows;
  }
  if (tiles == 0) return EDA_OK;
  p.nprobs = nprobs; p.K = K; p.Kpad = kpad_of(K); p.N = N; p.relu = relu; p.ln = layer_norm ? 1 : 0;
  p.gamma = ln_gamma; p.beta = ln_beta; p.eps = ln_eps;
  p.tmem_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : N <= 256 ? 256u : 512u;
  p.stage_bytes = (uint32_t)(kABytes + kKBlock * N * 4);
  const size_t smem = (size_t)kStages * p.stage_bytes;
  EDA_CUDA_TRY(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "linear smem attr");
  linear_kernel<<<tiles, kThreads, smem, as_stream(stream)>>>(p);
  return check_launch("linear_kernel");
}

}  // extern "C"
