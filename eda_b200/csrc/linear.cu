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
// CTA = one 128-row tile, 10 warps with fixed roles, K streamed in 32-wide blocks through a 3-stage ring:
//   warps 0-7  A staging: coalesced 16-byte cp.async copies (8 lanes = one 128-byte row segment) into the
//              K-major SWIZZLE_128B tile the tensor core reads, one block ahead of the maths; then every
//              thread adds "+ pos" to and rounds (cvt.rna.tf32) the chunks it copied, in place, and
//              arrives on the stage's `a_ready` mbarrier.  Later: the epilogue.
//   warp 8     one thread issues tcgen05.mma kind::tf32 (M = 128, N <= 256 per instruction, N = 288 ->
//              2 x 144; fp32 accumulator 128 x N in tensor memory) and commits each stage back to `empty`
//   warp 9     one thread streams the pre-packed weight blocks (eda_linear_pack) with 1-D TMA bulk copies
// Measured on B200 (scripts/lin_ts.py): a tf32 MMA of this shape occupies the tensor pipe ~100 cycles, so
// the 72 MMAs of a 128x288x288 tile are the floor (~7.5k cycles); staging runs in their shadow.
// Epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (its hardware quadrant) and the column half w / 4, so
// two warps per scheduler hide each other's latency.  Results go through a padded shared-memory tile (the
// ring is free by then) and leave with coalesced 16-byte stores; the residual tile arrives by coalesced
// cp.async.  LayerNorm: per-row (sum, sum of squares) of the two column halves meet in shared memory and the
// normalisation is applied on the way out.
#include "umma.cuh"

namespace eda {
namespace {

constexpr int kRows = 128;
constexpr int kWorkers = 256;   // warps 0-7: staging + epilogue
constexpr int kThreads = 320;   // + warp 8 MMA issue, warp 9 weight producer
constexpr int kStages = 3;
constexpr int kKBlock = 32;                      // 32 fp32 = one 128-byte swizzle row
constexpr int kTileBytes = kRows * kKBlock * 4;  // 16 KB
constexpr int kABytes = 2 * kTileBytes;          // A tile + P ("+ pos") tile
constexpr int kMaxN = 320;
constexpr int kMaxProbs = 3;
constexpr int kPrefetch = 1;  // blocks the staging warps run ahead (the slot it needs was released a whole
                              // iteration ago, so staging never waits for the MMAs it has just enabled)

struct LinProblem {
  const float *x, *pos, *w, *bias, *residual;
  float *y;
  int rows, tile0;
  int tb, ldt;  // tb > 0: channel-major output, y[((row / tb) * N + col) * ldt + row % tb]
  int round_out;  // round the outputs to tf32 (they feed tensor-core operands of the attention kernel directly)
};

struct LinParams {
  LinProblem pr[kMaxProbs];
  int nprobs, K, Kpad, N, relu, ln;
  const float *gamma, *beta;
  float eps;
  uint32_t tmem_cols, stage_bytes;
  uint32_t drop_thresh, drop_seed;  // output dropout (after bias / ReLU, before residual): thresh 0 = off
  float drop_scale;                 // 1 / (1 - p)
};

__device__ __forceinline__ void named_bar_sync_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32x4(float4 a) {
  return make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk store shared -> global (16-byte aligned, size multiple of 16), bulk-group completion.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Development aid: phase timestamps (clock64) of CTA 0 of the most recent launch (eda_debug_timestamps):
// [0] entry, [1] setup done, [2] staging loop done, [3] accumulator complete, [16] epilogue done, [17] exit.
__device__ long long g_lin_ts[32];
#define LIN_TS(i) do { if (blockIdx.x == 0 && tid == 0) g_lin_ts[i] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const LinParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B atoms are addressed by absolute shared-memory address bits: align the ring to 1024 bytes
  unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_w[kStages], a_ready[kStages], empty[kStages], mma_done, res_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[kMaxN], s_gamma[kMaxN], s_beta[kMaxN];
  __shared__ float s_part[2][kRows], s_part2[2][kRows];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  LIN_TS(0);
  int pi = 0;
  while (pi + 1 < p.nprobs && (int)blockIdx.x >= p.pr[pi + 1].tile0) ++pi;
  const LinProblem &pr = p.pr[pi];
  const int tile = (int)blockIdx.x - pr.tile0;
  const int N = p.N, K = p.K;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&a_ready[s], kWorkers);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&mma_done, 1);
    mbar_init(&res_bar, 1);
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
  LIN_TS(1);

  if (warp == 9) {
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
  } else if (warp == 8) {
    // ---------------- MMA issuer -----------------------------------------------------------------
    if (lane == 0) {
      // N split into MMA-sized pieces (multiples of 16, <= 256)
      const int n_a = N <= 256 ? N : ((N / 2 + 15) / 16) * 16;
      const int n_b = N - n_a;
      const uint32_t idesc_a = umma::idesc_tf32(kRows, n_a);
      const uint32_t idesc_b = umma::idesc_tf32(kRows, n_b > 0 ? n_b : 16);
      const uint32_t lbo_w = (uint32_t)N * 16u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % kStages;
        const int kcnt = min(kKBlock, p.Kpad - kb * kKBlock);
        const uint32_t par = (kb / kStages) & 1;
        mbar_wait(&a_ready[slot], par);
        mbar_wait(&full_w[slot], par);
        umma::fence_after_thread_sync();
        const uint32_t abase = smem_u32(smem_raw + (size_t)slot * p.stage_bytes);
        const uint32_t wbase = abase + kABytes;
        for (int ks = 0; ks < kcnt / 8; ++ks) {
          // A: K-major SWIZZLE_128B (rows of 128 B, 8-row atoms 1024 B apart); +32 B per 8-wide k step
          const uint64_t adesc = umma::smem_desc_swizzled(abase + (uint32_t)ks * 32u, 16u, 1024u, 2u);
          const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
          const uint64_t b0 = umma::smem_desc_kmajor_noswizzle(wbase + (uint32_t)ks * 2u * lbo_w, lbo_w, 128u);
          umma::mma_tf32_ss(tbase, adesc, b0, idesc_a, acc);
          if (n_b > 0) {
            const uint64_t b1 =
                umma::smem_desc_kmajor_noswizzle(wbase + (uint32_t)ks * 2u * lbo_w + (uint32_t)n_a * 16u, lbo_w, 128u);
            umma::mma_tf32_ss(tbase + (uint32_t)n_a, adesc, b1, idesc_b, acc);
          }
        }
        umma::mma_commit(&empty[slot]);
        if (kb == nkb - 1) umma::mma_commit(&mma_done);
      }
    }
  } else {
    // ---------------- staging warps ---------------------------------------------------------------
    const long long row0 = (long long)tile * kRows;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(pr.x) & 15) == 0 &&
                     (!pr.pos || (reinterpret_cast<uintptr_t>(pr.pos) & 15) == 0);
    // chunk ownership: 16-byte chunk j = tid % 8 of rows r_i = i * 32 + tid / 8, i = 0..3
    constexpr int kOwn = kRows * 8 / kWorkers;  // chunks per thread per block
    constexpr int kRowStep = kWorkers / 8;
    const int cj = tid & 7, cr0 = tid >> 3;

    auto issue_block = [&](int kb) {
      if (vec && kb < nkb) {
        const int slot = kb % kStages;
        mbar_wait(&empty[slot], ((kb / kStages) & 1) ^ 1);  // the MMAs that read this slot have finished
        unsigned char *sA = smem_raw + (size_t)slot * p.stage_bytes;
        unsigned char *sP = sA + kTileBytes;
        const int k = kb * kKBlock + cj * 4;
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const bool in = (row0 + r < pr.rows) && k < K;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
          const size_t g = (size_t)(row0 + r) * K + k;
          umma::cp_async16(sA + off, in ? pr.x + g : pr.x, in ? 16u : 0u);
          if (pr.pos) umma::cp_async16(sP + off, in ? pr.pos + g : pr.pos, in ? 16u : 0u);
        }
      }
      umma::cp_async_commit();
    };
    for (int i = 0; i < kPrefetch; ++i) issue_block(i);

    for (int kb = 0; kb < nkb; ++kb) {
      const int slot = kb % kStages;
      issue_block(kb + kPrefetch);
      umma::cp_async_wait<kPrefetch>();  // block kb has landed (this thread's copies)
      unsigned char *sA = smem_raw + (size_t)slot * p.stage_bytes;
      if (vec) {
        // the chunks this thread copied: "+ pos", round to tf32 (round-to-nearest), in place
        const unsigned char *sP = sA + kTileBytes;
        float4 v[kOwn];
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
          v[i] = *reinterpret_cast<const float4 *>(sA + off);
          if (pr.pos) {
            const float4 q = *reinterpret_cast<const float4 *>(sP + off);
            v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
          }
        }
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
          *reinterpret_cast<float4 *>(sA + off) = tf32x4(v[i]);
        }
      } else {
        // rows that are not 16-byte aligned (K = 3 or 6: xyz / box inputs of the position embedding): thread = row
        mbar_wait(&empty[slot], ((kb / kStages) & 1) ^ 1);
        const long long row = row0 + tid;
        const bool valid = row < pr.rows;
        const float *xrow = pr.x + row * K;
        const float *prow = pr.pos ? pr.pos + row * K : nullptr;
        const int nch = tid < kRows ? (min(kKBlock, p.Kpad - kb * kKBlock) >> 2) : 0;
        for (int c = 0; c < nch; ++c) {
          const int k = kb * kKBlock + c * 4;
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[e] = 0.f;
            if (valid && k + e < K) f[e] = __ldg(xrow + k + e) + (prow ? __ldg(prow + k + e) : 0.f);
          }
          *reinterpret_cast<float4 *>(sA + (uint32_t)tid * 128u + (uint32_t)((c ^ (tid & 7)) << 4)) =
              make_float4(to_tf32(f[0]), to_tf32(f[1]), to_tf32(f[2]), to_tf32(f[3]));
        }
      }
      umma::fence_proxy_async_smem();
      mbar_arrive(&a_ready[slot]);
    }
    LIN_TS(2);
    mbar_wait(&mma_done, 0);
    umma::fence_after_thread_sync();
    __syncwarp();
    LIN_TS(3);

    // ---------------- epilogue: thread = (row, column half) ------------------------------------------------
    // All MMAs and weight copies are complete: the ring is reused as the output tile [128][N + 4].
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const long long row = row0 + r;
    const bool valid = row < pr.rows;
    const int NH = ((N / 2 + 15) / 16) * 16;  // columns per half (multiple of 16)
    const int cbeg = min(N, half * NH), cend = min(N, cbeg + NH);
    const uint32_t trow = umma::tmem_addr(tbase, (uint32_t)(q * 32), 0);
    const int pitch = N + 4;
    float *tile_s = reinterpret_cast<float *>(smem_raw);
    float *srow = tile_s + (size_t)r * pitch;
    const int nvalid = (int)min((long long)kRows, (long long)pr.rows - row0);
    const int n4 = N >> 2;
    if (p.ln && pr.residual) {
      // residual tile -> shared memory, coalesced 16-byte cp.async (a warp moves one row segment at a time)
      for (int rr = warp; rr < nvalid; rr += kWorkers / 32)
        for (int c4 = lane; c4 < n4; c4 += 32)
          umma::cp_async16(tile_s + (size_t)rr * pitch + c4 * 4, pr.residual + (row0 + rr) * N + c4 * 4, 16u);
      umma::cp_async_commit();
      umma::cp_async_wait<0>();
      named_bar_sync_workers();
    }
    LIN_TS(8);
    float sum = 0.f, sumsq = 0.f;
    const bool direct_t = pr.tb > 0;  // channel-major output goes straight to global (already coalesced)
    long long bb = 0, rr0 = 0;
    if (direct_t) { bb = row / pr.tb; rr0 = row - bb * pr.tb; }
    auto process = [&](const uint32_t (&u)[16], int c0) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int c = c0 + q4 * 4;
        const float4 b4 = *reinterpret_cast<const float4 *>(s_bias + c);
        float4 o;
        o.x = __uint_as_float(u[q4 * 4 + 0]) + b4.x;
        o.y = __uint_as_float(u[q4 * 4 + 1]) + b4.y;
        o.z = __uint_as_float(u[q4 * 4 + 2]) + b4.z;
        o.w = __uint_as_float(u[q4 * 4 + 3]) + b4.w;
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (p.drop_thresh) {  // nn.Dropout on this block's output: element (row, column) of problem pi
          const uint32_t ra = (uint32_t)row * (uint32_t)kMaxProbs + (uint32_t)pi;
          o.x = dropout_keep(p.drop_seed, ra, (uint32_t)c + 0u, p.drop_thresh) ? o.x * p.drop_scale : 0.f;
          o.y = dropout_keep(p.drop_seed, ra, (uint32_t)c + 1u, p.drop_thresh) ? o.y * p.drop_scale : 0.f;
          o.z = dropout_keep(p.drop_seed, ra, (uint32_t)c + 2u, p.drop_thresh) ? o.z * p.drop_scale : 0.f;
          o.w = dropout_keep(p.drop_seed, ra, (uint32_t)c + 3u, p.drop_thresh) ? o.w * p.drop_scale : 0.f;
        }
        if (pr.round_out) o = tf32x4(o);
        if (p.ln) {
          if (pr.residual && valid) {
            const float4 r4 = *reinterpret_cast<const float4 *>(srow + c);
            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
          }
          sum += (o.x + o.y) + (o.z + o.w);
          sumsq = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, sumsq))));
        }
        if (direct_t) {
          if (valid) {
            float *yt = pr.y + (bb * N + c) * (long long)pr.ldt + rr0;
            yt[0] = o.x; yt[pr.ldt] = o.y; yt[2 * (long long)pr.ldt] = o.z; yt[3 * (long long)pr.ldt] = o.w;
          }
        } else {
          *reinterpret_cast<float4 *>(srow + c) = o;
        }
      }
    };
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t u[16];
      umma::tmem_ld16(trow + (uint32_t)c0, u);
      umma::tmem_ld_wait();
      process(u, c0);
    }
    LIN_TS(9);
    if (p.ln) {
      // per-row (sum, sum of squares) of the two column halves; combined by the copy-out below.
      // var = E[x^2] - mean^2 in fp32: relative error ~1e-7 (1 + mean^2/var), harmless for these activations.
      s_part[half][r] = sum;
      s_part2[half][r] = sumsq;
    }
    LIN_TS(10);
    if (!direct_t) {
      named_bar_sync_workers();  // the tile (and the row statistics) are complete
      // coalesced copy-out (a warp moves one row at a time), LayerNorm applied on the way
      const float invN = 1.0f / (float)N;
      for (int rr = warp; rr < nvalid; rr += kWorkers / 32) {
        float mean = 0.f, rstd = 1.f;
        if (p.ln) {
          const float m1 = (s_part[0][rr] + s_part[1][rr]) * invN;
          const float m2 = (s_part2[0][rr] + s_part2[1][rr]) * invN;
          mean = m1;
          rstd = 1.0f / sqrtf(fmaxf(m2 - m1 * m1, 0.f) + p.eps);
        }
        const float4 *src = reinterpret_cast<const float4 *>(tile_s + (size_t)rr * pitch);
        float4 *dst = reinterpret_cast<float4 *>(pr.y + (row0 + rr) * N);
        for (int c4 = lane; c4 < n4; c4 += 32) {
          float4 o = src[c4];
          if (p.ln) {
            const float4 g4 = *reinterpret_cast<const float4 *>(s_gamma + c4 * 4);
            const float4 e4 = *reinterpret_cast<const float4 *>(s_beta + c4 * 4);
            o.x = (o.x - mean) * rstd * g4.x + e4.x;
            o.y = (o.y - mean) * rstd * g4.y + e4.y;
            o.z = (o.z - mean) * rstd * g4.z + e4.z;
            o.w = (o.w - mean) * rstd * g4.w + e4.w;
          }
          dst[c4] = o;
        }
      }
    }
    LIN_TS(16);
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, p.tmem_cols);
  LIN_TS(17);
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

// keep-mask (1.0 / 0.0) of the dropout decisions above, for the backward pass: out[a * cols + b]
__global__ void dropout_mask_kernel(uint32_t seed, uint32_t thresh, long long rows, int cols, uint32_t a_mul, uint32_t a_add,
                                    float *__restrict__ out) {
  const long long total = rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long a = e / cols;
    const int b = (int)(e - a * cols);
    out[e] = dropout_keep(seed, (uint32_t)a * a_mul + a_add, (uint32_t)b, thresh) ? 1.f : 0.f;
  }
}

inline int kpad_of(int K) { return (K + 7) & ~7; }
inline bool lin_supported(int N, int K) { return N >= 16 && N <= kMaxN && (N & 15) == 0 && K >= 1 && K <= 4096; }

}  // namespace
}  // namespace eda

extern "C" {

int eda_debug_timestamps(long long *host_out, int n) {
  if (!host_out || n < 0 || n > 32) return EDA_ERR_INVALID_ARGUMENT;
  EDA_CUDA_TRY(cudaMemcpyFromSymbol(host_out, eda::g_lin_ts, sizeof(long long) * n), "debug timestamps");
  return EDA_OK;
}

int eda_dropout_mask(unsigned int seed, float p, long long rows, int cols, unsigned int a_mul, unsigned int a_add,
                     float *out, void *stream) {
  using namespace eda;
  if (rows < 0 || cols < 0 || p < 0.f || p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0 || cols == 0) return EDA_OK;
  if (!out) return EDA_ERR_INVALID_ARGUMENT;
  const long long total = rows * cols;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  dropout_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(seed, dropout_thresh(p), rows, cols, a_mul, a_add, out);
  return check_launch("dropout_mask_kernel");
}

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
                       const float *ln_beta, float ln_eps, int layer_norm, float dropout_p, unsigned int dropout_seed,
                       void *stream) {
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
    if (q.y_batch_rows < 0 || (q.y_batch_rows > 0 && (q.y_ld < q.y_batch_rows || layer_norm)))
      return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].x = q.x; p.pr[i].pos = q.pos; p.pr[i].w = q.w_packed; p.pr[i].bias = q.bias;
    p.pr[i].residual = q.residual; p.pr[i].y = q.y; p.pr[i].rows = q.rows; p.pr[i].tile0 = tiles;
    p.pr[i].tb = q.y_batch_rows; p.pr[i].ldt = q.y_ld; p.pr[i].round_out = q.round_tf32;
    tiles += (q.rows + kRows - 1) / kRows;
  }
  if (tiles == 0) return EDA_OK;
  p.nprobs = nprobs; p.K = K; p.Kpad = kpad_of(K); p.N = N; p.relu = relu; p.ln = layer_norm ? 1 : 0;
  p.gamma = ln_gamma; p.beta = ln_beta; p.eps = ln_eps;
  if (dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  p.drop_thresh = dropout_thresh(dropout_p); p.drop_seed = dropout_seed; p.drop_scale = 1.0f / (1.0f - dropout_p);
  p.tmem_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : N <= 256 ? 256u : 512u;
  // stage = A tile + P tile + W block, padded to 1024 bytes (SWIZZLE_128B atoms must stay 1024-aligned)
  p.stage_bytes = (uint32_t)((kABytes + kKBlock * N * 4 + 1023) & ~1023);
  size_t smem = (size_t)kStages * p.stage_bytes;
  const size_t out_tile = (size_t)kRows * (N + 4) * sizeof(float);
  if (smem < out_tile) smem = out_tile;
  smem += 1024;  // alignment slack
  static size_t smem_set = 0;  // one process per GPU: the attribute is raised once per size increase
  if (smem > smem_set) {
    EDA_CUDA_TRY(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "linear smem attr");
    smem_set = smem;
  }
  linear_kernel<<<tiles, kThreads, smem, as_stream(stream)>>>(p);
  return check_launch("linear_kernel");
}

}  // extern "C"
