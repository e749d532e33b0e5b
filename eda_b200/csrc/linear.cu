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
// CTA = one 128-row tile, 18 warps with fixed roles, K streamed in 32-wide blocks through a 3-stage ring:
//   warps 0-15 A staging: coalesced 16-byte cp.async copies (8 lanes = one 128-byte row segment) into the
//              K-major SWIZZLE_128B tile the tensor core reads, one block ahead of the maths; then every
//              thread adds "+ pos" to and rounds (cvt.rna.tf32) the chunks it copied, in place, and
//              arrives on the stage's `a_ready` mbarrier.  Later: the epilogue.
//   warp 16    one thread issues tcgen05.mma kind::tf32 (M = 128, N <= 256 per instruction, N = 288 ->
//              2 x 144; fp32 accumulator 128 x N in tensor memory) and commits each stage back to `empty`
//   warp 17    one thread streams the pre-packed weight blocks (eda_linear_pack) with 1-D TMA bulk copies
// Measured on B200 (scripts/lin_ts.py): a tf32 MMA of this shape occupies the tensor pipe ~100 cycles, so
// the 72 MMAs of a 128x288x288 tile are the floor (~7.5k cycles); staging runs in their shadow.
// Epilogue, two phases over a padded shared-memory tile (the ring is free by then).  Phase 1, thread = (row,
// column quarter): warp w reads TMEM lanes 32 (w % 4).. (its hardware quadrant), four warps per scheduler hide
// each other's TMEM latency: bias, ReLU, dropout, rounding.  Phase 2, warp = row: residual read with coalesced
// 16-byte loads, LayerNorm statistics by warp shuffle, coalesced 16-byte stores.
#include "umma.cuh"
#include "tensor_map.cuh"
#include "wgrad_tc.h"

namespace eda {
namespace {

constexpr int kRows = 128;
constexpr int kWorkers = 512;   // warps 0-15: staging + epilogue
constexpr int kWorkerWarps = kWorkers / 32;
constexpr int kThreads = 576;   // + warp 16 MMA issue, warp 17 weight producer
constexpr int kStages = 4;  // ring slots allocated at most; p.nstages (3 or 4) are used
constexpr int kKBlock = 32;                      // 32 fp32 = one 128-byte swizzle row
constexpr int kTileBytes = kRows * kKBlock * 4;  // 16 KB
constexpr int kABytes = kTileBytes;              // A tile ("+ pos" is added from global memory during the fix-up pass)
constexpr int kMaxN = 320;
constexpr int kMaxProbs = 3;
constexpr int kPrefetch = 1;  // blocks the staging warps run ahead (the slot it needs was released a whole
                              // iteration ago, so staging never waits for the MMAs it has just enabled)

struct LinProblem {
  int tma;  // A tiles arrive by 2-D TMA (tensor map LinParams::tmap[problem]) instead of per-thread cp.async
  const float *x, *pos, *w, *bias, *residual;
  float *y;
  float *pre;  // optional (rows, N): the LayerNorm INPUT (residual + product), saved for the backward pass
  int rows, tile0;
  int ldy;      // row stride of the row-major output (>= N): y may be a column block of a wider matrix
  int tb, ldt;  // tb > 0: channel-major output, y[((row / tb) * N + col) * ldt + row % tb]
  int round_out;  // round the outputs to tf32 (they feed tensor-core operands of the attention kernel directly)
};

struct LinParams {
  // one tensor map per problem over x as a (rows, K) fp32 matrix: box = 32 columns x 128 rows, SWIZZLE_128B — the
  // TMA engine writes exactly the K-major swizzled tile the tensor core reads (chunk j of row r at j ^ (r & 7)), rows /
  // columns beyond the matrix are zero-filled
  alignas(64) CUtensorMap tmap[kMaxProbs];
  LinProblem pr[kMaxProbs];
  int nprobs, K, Kpad, N, relu, ln;
  int S, NS;  // the N columns are split over S CTAs of NS columns each (a cluster when LayerNorm needs whole rows)
  const float *gamma, *beta;
  float eps;
  uint32_t tmem_cols, stage_bytes;
  int nstages, prefetch;  // ring depth in use and how many blocks the staging warps run ahead (nstages - 2)
  int sub;                // 32-wide K sub-blocks per ring stage (1..3; > 1 only when every problem's A tiles come by TMA): a
                          // stage then carries 8 - 24 MMAs per commit instead of 4 - 8 — the tensor pipe needs ~400 cycles
                          // to fill and drain around every commit (eda_selftest_umma_rate), so few fat stages beat many
                          // thin ones
  uint32_t a_bytes;       // sub * 16 KB: offset of the W block inside a stage
  int ts_hack;            // timing experiments only (EDA_LIN_TSHACK bits: 2 no A traffic, 4 no W traffic, 8 no fix-up pass,
                          // 16 MMAs issued without waiting for operands — results are garbage)
  uint32_t drop_thresh, drop_seed;  // output dropout (after bias / ReLU, before residual): thresh 0 = off
  const uint32_t *seed_epoch;       // optional device word added to drop_seed (eda_dropout_set_epoch)
  float drop_scale;                 // 1 / (1 - p)
};

__device__ __forceinline__ void named_bar_sync_workers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32x4(float4 a) {
  return make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
}
__device__ __forceinline__ float2 ld_shared_cluster_f32x2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
  return v;
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
__device__ long long g_lin_ts[128];
#define LIN_TS(i) do { if (blockIdx.x == 0 && tid == 0) g_lin_ts[i] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const __grid_constant__ LinParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B atoms are addressed by absolute shared-memory address bits: align the ring to 1024 bytes
  unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_w[kStages], full_a[kStages], a_ready[kStages], empty[kStages], mma_done, res_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[kMaxN], s_gamma[kMaxN], s_beta[kMaxN];
  __shared__ __align__(8) float2 s_stat[kRows];  // per-row (sum, sum of squares) of this CTA's column slice

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();  // == warp, known warp-uniform to the compiler
  LIN_TS(0);
  const int gtile = (int)blockIdx.x / p.S, split = (int)blockIdx.x - gtile * p.S;
  int pi = 0;
  while (pi + 1 < p.nprobs && gtile >= p.pr[pi + 1].tile0) ++pi;
  const LinProblem &pr = p.pr[pi];
  const int tile = gtile - pr.tile0;
  const int Nf = p.N, N = p.NS, n0 = split * p.NS, K = p.K;  // N: this CTA's slice [n0, n0 + N) of the Nf columns
  const bool clustered = p.ln && p.S > 1;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_a[s], 1);
      mbar_init(&a_ready[s], kWorkerWarps);  // one arrival per staging warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(&mma_done, 1);
    mbar_init(&res_bar, 1);
    mbar_fence_init_cluster();
  }
  // everything above is private set-up (tensor-memory allocation, barrier initialisation); from here on memory that the
  // predecessor in the stream may have written is read (bias / shift vectors can come straight out of a BatchNorm
  // finalisation, activations and packed weights anyway)
  pdl_launch_dependents();
  pdl_wait();
  for (int i = tid; i < N; i += kThreads) {
    s_bias[i] = pr.bias ? __ldg(pr.bias + n0 + i) : 0.f;
    s_gamma[i] = (p.ln && p.gamma) ? __ldg(p.gamma + n0 + i) : 1.f;
    s_beta[i] = (p.ln && p.beta) ? __ldg(p.beta + n0 + i) : 0.f;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  const int KB = kKBlock * p.sub;                 // K columns per ring stage
  const int nkb = (p.Kpad + KB - 1) / KB;
  LIN_TS(1);

  if (warp_u == kWorkerWarps + 1) {
    // ---------------- weight producer ----------------------------------------------------------
    if (lane == 0) {
      // ring position as running counters: kb % nstages / kb / nstages by a run-time divisor cost a few hundred cycles
      // of dependent scalar code per K block on these single-thread control paths
      int slot = 0;
      uint32_t round = 0;
      for (int kb = 0; kb < nkb; ++kb, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, round ^= (slot == 0) ? 1u : 0u) {
        const int kcnt = min(KB, p.Kpad - kb * KB);
        const uint32_t bytes = (uint32_t)kcnt * (uint32_t)N * 4u;
        mbar_wait(&empty[slot], round ^ 1u);
        if (pr.tma) {  // A: one box of 32 columns x 128 rows per sub-block, out-of-range parts zero-filled
          if (p.ts_hack & 2) {
            mbar_arrive(&full_a[slot]);  // timing experiment: no A traffic
          } else {
            const int nsub = (kcnt + kKBlock - 1) / kKBlock;
            mbar_arrive_expect_tx(&full_a[slot], (uint32_t)(nsub * kTileBytes));
            for (int j = 0; j < nsub; ++j)
              tma_load_2d(smem_raw + (size_t)slot * p.stage_bytes + (size_t)j * kTileBytes, &p.tmap[pi],
                          kb * KB + j * kKBlock, tile * kRows, &full_a[slot]);
          }
        }
        if (p.ts_hack & 4) { mbar_arrive(&full_w[slot]); continue; }  // timing experiment: no W traffic
        mbar_arrive_expect_tx(&full_w[slot], bytes);
        unsigned char *wdst = smem_raw + (size_t)slot * p.stage_bytes + p.a_bytes;
        const float *wsrc = pr.w + (size_t)kb * KB * Nf;  // packed: [16-byte chunk of K][Nf columns] float4, K ascending
        if (p.S == 1) {
          bulk_g2s(wdst, wsrc, bytes, &full_w[slot]);
        } else {
          // this slice's NS columns of every chunk row: contiguous pieces of NS * 16 bytes
          for (int c = 0; c < kcnt / 4; ++c)
            bulk_g2s(wdst + (size_t)c * N * 16, wsrc + ((size_t)c * Nf + n0) * 4, (uint32_t)N * 16u, &full_w[slot]);
        }
      }
    }
    if (clustered) cluster_sync_all();  // matches the workers' statistics exchange
  } else if (warp_u == kWorkerWarps) {
    // ---------------- MMA issuer -----------------------------------------------------------------
    // The WHOLE warp runs this loop, converged, with warp-uniform values (the role branch is on a shuffled warp index,
    // so the compiler knows): descriptors stay in uniform registers and the four K steps of a 32-wide sub-block go out
    // from one statement (umma::mma4_tf32_ss_w: one elected lane, two adds per MMA).  Written as `if (lane == 0)` with
    // per-MMA descriptor arithmetic the same loop cost 160 - 330 cycles per MMA in scalar issue code (vector -> uniform
    // register moves in an elect / branch "waterfall" around every tcgen05.mma) against 72 for the MMA itself
    // (eda_selftest_umma_rate, scripts/umma_rate.py).
    {
      // N split into MMA-sized pieces (multiples of 16, <= 256)
      const int n_a = N <= 256 ? N : ((N / 2 + 15) / 16) * 16;
      const int n_b = N - n_a;
      const uint32_t idesc_a = umma::idesc_tf32(kRows, n_a);
      const uint32_t idesc_b = umma::idesc_tf32(kRows, n_b > 0 ? n_b : 16);
      const uint32_t lbo_w = (uint32_t)N * 16u;
      const uint32_t ring0 = smem_u32(smem_raw);
      const uint64_t ad0 = umma::smem_desc_swizzled(ring0, 16u, 1024u, 2u);
      const uint64_t bd0 = umma::smem_desc_kmajor_noswizzle(ring0 + p.a_bytes, lbo_w, 128u);
      const uint32_t a_hi = umma::desc_hi(ad0), b_hi = umma::desc_hi(bd0);
      // descriptor low words count 16-byte units: a K step is +32 bytes inside A's swizzle atom and two chunk rows of B
      const uint32_t a_step = 32u >> 4, b_step = (2u * lbo_w) >> 4, b_half = ((uint32_t)n_a * 16u) >> 4;
      const uint32_t a_sub = (uint32_t)kTileBytes >> 4, stage_step = p.stage_bytes >> 4;
      int slot = 0;
      uint32_t par = 0;
      for (int kb = 0; kb < nkb; ++kb, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
        const int kcnt = min(KB, p.Kpad - kb * KB);
        if (!(p.ts_hack & 16)) mbar_wait(&a_ready[slot], par);  // (timing experiment 16: issue without waiting)
        if (blockIdx.x == 0 && lane == 0 && kb < 10) g_lin_ts[32 + 3 * kb] = clock64();  // [32 + 3 kb] A block kb staged
        if (!(p.ts_hack & 16)) mbar_wait(&full_w[slot], par);
        if (blockIdx.x == 0 && lane == 0 && kb < 10) g_lin_ts[33 + 3 * kb] = clock64();  // [33 + 3 kb] W block kb landed
        umma::fence_after_thread_sync();
        uint32_t a_lo = umma::desc_lo(ad0) + (uint32_t)slot * stage_step;
        uint32_t b_lo = umma::desc_lo(bd0) + (uint32_t)slot * stage_step;
        for (int k0 = 0; k0 < kcnt; k0 += kKBlock, a_lo += a_sub, b_lo += 4u * b_step) {  // sub-blocks of the stage
          const int nks = min(kKBlock, kcnt - k0) / 8;
          const uint32_t acc = (kb > 0 || k0 > 0) ? 1u : 0u;
          if (nks == 4) {
            umma::mma4_tf32_ss_w(tbase, a_lo, a_hi, a_step, b_lo, b_hi, b_step, idesc_a, acc);
            if (n_b > 0)
              umma::mma4_tf32_ss_w(tbase + (uint32_t)n_a, a_lo, a_hi, a_step, b_lo + b_half, b_hi, b_step, idesc_b, acc);
          } else {  // K tail (Kpad % 32 != 0: the 3- / 6- / 8-wide inputs of the position embeddings)
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)ks * a_step);
              const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)ks * b_step);
              umma::mma_tf32_ss_w(tbase, ad, bd, idesc_a, (acc || ks > 0) ? 1u : 0u);
              if (n_b > 0) umma::mma_tf32_ss_w(tbase + (uint32_t)n_a, ad, bd + b_half, idesc_b, (acc || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma::mma_commit_w(&empty[slot]);
        if (blockIdx.x == 0 && lane == 0 && kb < 10) g_lin_ts[34 + 3 * kb] = clock64();  // [34 + 3 kb] MMAs of block kb issued
        if (kb == nkb - 1) umma::mma_commit_w(&mma_done);
      }
    }
    if (clustered) cluster_sync_all();  // matches the workers' statistics exchange
  } else {
    // ---------------- staging warps ---------------------------------------------------------------
    const long long row0 = (long long)tile * kRows;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(pr.x) & 15) == 0 &&
                     (!pr.pos || (reinterpret_cast<uintptr_t>(pr.pos) & 15) == 0);
    // chunk ownership: 16-byte chunk j = tid % 8 of rows r_i = i * 32 + tid / 8, i = 0..3
    constexpr int kOwn = kRows * 8 / kWorkers;  // chunks per thread per block
    constexpr int kRowStep = kWorkers / 8;
    const int cj = tid & 7, cr0 = tid >> 3;

    const bool tma = pr.tma != 0;
    auto issue_block = [&](int kb) {
      if (tma) return;  // the producer warp's TMA copies fill the A tiles
      if (vec && kb < nkb) {
        const int slot = p.nstages == 4 ? (kb & 3) : kb % 3;
        const int rnd = p.nstages == 4 ? (kb >> 2) : kb / 3;
        if (lane == 0) mbar_wait(&empty[slot], (uint32_t)((rnd & 1) ^ 1));  // the MMAs that read this slot are done
        __syncwarp();
        unsigned char *sA = smem_raw + (size_t)slot * p.stage_bytes;
        const int k = kb * kKBlock + cj * 4;
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const bool in = (row0 + r < pr.rows) && k < K;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
          const size_t g = (size_t)(row0 + r) * K + k;
          umma::cp_async16(sA + off, in ? pr.x + g : pr.x, in ? 16u : 0u);
        }
      }
      umma::cp_async_commit();
    };
    for (int i = 0; i < p.prefetch; ++i) issue_block(i);

    int slot = 0;
    uint32_t spar = 0;
    for (int kb = 0; kb < nkb; ++kb, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, spar ^= (slot == 0) ? 1u : 0u) {
#define STG_TS(j) do { if (blockIdx.x == 0 && tid == 0 && kb < 9) g_lin_ts[64 + 6 * kb + (j)] = clock64(); } while (0)
      STG_TS(0);
      if (tma) {
        // ---- TMA-fed stage (one to three 32-wide sub-blocks): wait for the boxes, then "+ pos" and tf32 rounding of
        // the chunks this thread owns in every sub-block, in place
        const int kcnt = min(KB, p.Kpad - kb * KB);
        const int nsub = (kcnt + kKBlock - 1) / kKBlock;
        float4 pq[3][kOwn];
        if (pr.pos) {  // issued before the wait: their latency hides behind the TMA transfer
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < nsub) {
              const int k = kb * KB + j * kKBlock + cj * 4;
#pragma unroll
              for (int i = 0; i < kOwn; ++i) {
                const int r = i * kRowStep + cr0;
                const bool in = (row0 + r < pr.rows) && k < K;
                pq[j][i] = in ? __ldg(reinterpret_cast<const float4 *>(pr.pos + (size_t)(row0 + r) * K + k))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
        }
        STG_TS(1);
        mbar_wait(&full_a[slot], spar);
        STG_TS(2);
        if (!(p.ts_hack & 8)) {
          unsigned char *sA0 = smem_raw + (size_t)slot * p.stage_bytes;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < nsub) {
              unsigned char *sA = sA0 + (size_t)j * kTileBytes;
              float4 v[kOwn];
#pragma unroll
              for (int i = 0; i < kOwn; ++i) {
                const int r = i * kRowStep + cr0;
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
                v[i] = *reinterpret_cast<const float4 *>(sA + off);
                if (pr.pos) {
                  v[i].x += pq[j][i].x; v[i].y += pq[j][i].y; v[i].z += pq[j][i].z; v[i].w += pq[j][i].w;
                }
              }
#pragma unroll
              for (int i = 0; i < kOwn; ++i) {
                const int r = i * kRowStep + cr0;
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
                *reinterpret_cast<float4 *>(sA + off) = tf32x4(v[i]);
              }
            }
          }
        }
        STG_TS(3);
        umma::fence_proxy_async_smem();
        STG_TS(4);
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[slot]);
        STG_TS(5);
        continue;
      }
      issue_block(kb + p.prefetch);
      STG_TS(1);
      // "+ pos" for block kb comes straight from global memory (the same 16-byte chunks this thread copied of x): the
      // loads are issued before the wait below, so their latency hides behind the cp.async of x
      float4 pq[kOwn];
      if (vec && pr.pos) {
        const int k = kb * kKBlock + cj * 4;
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const bool in = (row0 + r < pr.rows) && k < K;
          pq[i] = in ? __ldg(reinterpret_cast<const float4 *>(pr.pos + (size_t)(row0 + r) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (p.prefetch == 2) umma::cp_async_wait<2>(); else umma::cp_async_wait<1>();  // block kb has landed (own copies)
      STG_TS(2);
      unsigned char *sA = smem_raw + (size_t)slot * p.stage_bytes;
      if (vec && (p.ts_hack & 8)) {
        // timing experiment: no fix-up pass
      } else if (vec) {
        // the chunks this thread copied: "+ pos", round to tf32 (round-to-nearest), in place
        float4 v[kOwn];
#pragma unroll
        for (int i = 0; i < kOwn; ++i) {
          const int r = i * kRowStep + cr0;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cj ^ (r & 7)) << 4);
          v[i] = *reinterpret_cast<const float4 *>(sA + off);
          if (pr.pos) {
            v[i].x += pq[i].x; v[i].y += pq[i].y; v[i].z += pq[i].z; v[i].w += pq[i].w;
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
        mbar_wait(&empty[slot], spar ^ 1u);
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
      STG_TS(3);
      umma::fence_proxy_async_smem();
      STG_TS(4);
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[slot]);
      STG_TS(5);
    }
    LIN_TS(2);
    mbar_wait(&mma_done, 0);
    umma::fence_after_thread_sync();
    __syncwarp();
    LIN_TS(3);

    // ---------------- epilogue ---------------------------------------------------------------------------
    // All MMAs and weight copies are complete: the ring is reused as the output tile [128][N + 4].
    // Phase 1, thread = (row, column quarter): warp w reads TMEM lanes 32 (w % 4).. (its hardware quadrant) and
    // the 16-column chunks of quarter w / 4: bias, ReLU, dropout, rounding -> tile (or straight to the
    // channel-major output).  Phase 2, warp = row: residual (coalesced global reads) + LayerNorm with warp-shuffle
    // row statistics + coalesced 16-byte stores.
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const long long row = row0 + r;
    const bool valid = row < pr.rows;
    const int nchunks = N >> 4;
    const int cbeg = ((part * nchunks) / 4) << 4, cend = (((part + 1) * nchunks) / 4) << 4;
    const uint32_t trow = umma::tmem_addr(tbase, (uint32_t)(q * 32), 0);
    const int pitch = N + 4;
    float *tile_s = reinterpret_cast<float *>(smem_raw);
    float *srow = tile_s + (size_t)r * pitch;
    const int nvalid = (int)min((long long)kRows, (long long)pr.rows - row0);
    const int n4 = N >> 2;
    LIN_TS(8);
    const uint32_t dseed = p.drop_thresh ? effective_seed(p.drop_seed, p.seed_epoch) : 0u;
    const bool direct_t = pr.tb > 0;  // channel-major output goes straight to global (already coalesced)
    long long bb = 0, rr0 = 0;
    if (direct_t) { bb = row / pr.tb; rr0 = row - bb * pr.tb; }
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t u[16];
      umma::tmem_ld16(trow + (uint32_t)c0, u);
      umma::tmem_ld_wait();
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
          const uint32_t gc = (uint32_t)(n0 + c);
          o.x = dropout_keep(dseed, ra, gc + 0u, p.drop_thresh) ? o.x * p.drop_scale : 0.f;
          o.y = dropout_keep(dseed, ra, gc + 1u, p.drop_thresh) ? o.y * p.drop_scale : 0.f;
          o.z = dropout_keep(dseed, ra, gc + 2u, p.drop_thresh) ? o.z * p.drop_scale : 0.f;
          o.w = dropout_keep(dseed, ra, gc + 3u, p.drop_thresh) ? o.w * p.drop_scale : 0.f;
        }
        if (pr.round_out) o = tf32x4(o);
        if (direct_t) {
          if (valid) {
            float *yt = pr.y + (bb * Nf + n0 + c) * (long long)pr.ldt + rr0;
            yt[0] = o.x; yt[pr.ldt] = o.y; yt[2 * (long long)pr.ldt] = o.z; yt[3 * (long long)pr.ldt] = o.w;
          }
        } else {
          *reinterpret_cast<float4 *>(srow + c) = o;
        }
      }
    }
    LIN_TS(9);
    if (!direct_t && !p.ln) {
      // no LayerNorm: the tile already holds the final values.  One TMA bulk store per row (its N columns are contiguous
      // on both sides, 16-byte aligned) instead of 16-byte stores by warp = row, whose second pass over a 144-column row
      // used 4 of 32 lanes: 4.3k cycles of a 21k-cycle CTA.
      umma::fence_proxy_async_smem();  // the bulk copy reads the tile through the async proxy
      named_bar_sync_workers();        // the tile is complete
      LIN_TS(10);
      if (tid < nvalid) {
        const float *src = tile_s + (size_t)tid * pitch;
        bulk_s2g(pr.y + (row0 + tid) * (long long)pr.ldy + n0, src, (uint32_t)N * 4u);
        if (pr.pre) bulk_s2g(pr.pre + (row0 + tid) * Nf + n0, src, (uint32_t)N * 4u);
        bulk_commit();
        bulk_wait_read_all();  // shared memory may go once the copies have read it; the grid's completion publishes them
      }
      LIN_TS(16);
    } else if (!direct_t) {
      named_bar_sync_workers();  // the tile is complete
      LIN_TS(10);
      const float invN = 1.0f / (float)Nf;
      constexpr int kMaxJ = (kMaxN / 4 + 31) / 32;  // float4 columns per lane
      if (p.ln) {
        // pass A (warp = 8 rows, 4 at a time so 12 residual loads are in flight): residual added in place, per-row
        // (sum, sum of squares) of this CTA's columns -> s_stat
        constexpr int kRB = 4;
        for (int rb = warp * (kRows / kWorkerWarps); rb < (warp + 1) * (kRows / kWorkerWarps); rb += kRB) {
          float4 r4[kRB][kMaxJ];
#pragma unroll
          for (int i = 0; i < kRB; ++i)
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) {
              const int c4 = lane + j * 32;
              r4[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (pr.residual && rb + i < nvalid && c4 < n4)
                r4[i][j] = __ldg(reinterpret_cast<const float4 *>(pr.residual + (row0 + rb + i) * Nf + n0) + c4);
            }
          if (rb == warp * (kRows / kWorkerWarps)) LIN_TS(13);
#pragma unroll
          for (int i = 0; i < kRB; ++i) {
            float sum = 0.f, sumsq = 0.f;
            float4 *src = reinterpret_cast<float4 *>(tile_s + (size_t)(rb + i) * pitch);
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) {
              const int c4 = lane + j * 32;
              if (rb + i < nvalid && c4 < n4) {
                float4 o = src[c4];
                o.x += r4[i][j].x; o.y += r4[i][j].y; o.z += r4[i][j].z; o.w += r4[i][j].w;
                if (pr.residual) src[c4] = o;
                sum += (o.x + o.y) + (o.z + o.w);
                sumsq = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, sumsq))));
              }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
              sum += __shfl_xor_sync(0xffffffffu, sum, d);
              sumsq += __shfl_xor_sync(0xffffffffu, sumsq, d);
            }
            if (lane == 0) s_stat[rb + i] = make_float2(sum, sumsq);
          }
        }
        LIN_TS(11);
        if (clustered) cluster_sync_all(); else __syncwarp();
        LIN_TS(12);
      }
      // pass B: normalise with the row statistics (own, or all S column slices' through DSMEM) and store.
      // The statistics of all of this warp's rows are fetched first: a remote shared-memory read costs ~1k cycles and
      // eight of them in a row (one per loop iteration) were most of this pass.
      constexpr int kRowsPerWarp = kRows / kWorkerWarps;
      float w_mean[kRowsPerWarp], w_rstd[kRowsPerWarp];
      if (p.ln) {
        float2 st[kRowsPerWarp];
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
          const int rr = warp * kRowsPerWarp + i;
          st[i] = make_float2(0.f, 0.f);
          if (clustered) {
            if (lane < p.S) st[i] = ld_shared_cluster_f32x2(mapa_u32(smem_u32(&s_stat[rr]), (uint32_t)lane));
          } else {
            st[i] = s_stat[rr];
          }
        }
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
          float sum = st[i].x, sumsq = st[i].y;
          if (clustered) {
#pragma unroll
            for (int d = 2; d > 0; d >>= 1) {  // S <= 4
              sum += __shfl_xor_sync(0xffffffffu, sum, d);
              sumsq += __shfl_xor_sync(0xffffffffu, sumsq, d);
            }
            sum = __shfl_sync(0xffffffffu, sum, 0);
            sumsq = __shfl_sync(0xffffffffu, sumsq, 0);
          }
          // var = E[x^2] - mean^2 in fp32: relative error ~1e-7 (1 + mean^2/var), harmless for these activations
          w_mean[i] = sum * invN;
          w_rstd[i] = 1.0f / sqrtf(fmaxf(sumsq * invN - w_mean[i] * w_mean[i], 0.f) + p.eps);
        }
      }
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int rr = warp * kRowsPerWarp + i;
        if (rr >= nvalid) break;
        const float4 *src = reinterpret_cast<const float4 *>(tile_s + (size_t)rr * pitch);
        float4 *dst = reinterpret_cast<float4 *>(pr.y + (row0 + rr) * (long long)pr.ldy + n0);
        const float mean = p.ln ? w_mean[i] : 0.f, rstd = p.ln ? w_rstd[i] : 1.f;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
          const int c4 = lane + j * 32;
          if (c4 < n4) {
            float4 o = src[c4];
            if (pr.pre) reinterpret_cast<float4 *>(pr.pre + (row0 + rr) * Nf + n0)[c4] = o;
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
    } else if (clustered) {
      cluster_sync_all();  // unreachable (LayerNorm never combines with channel-major output); keeps arrival counts equal
    }
    LIN_TS(16);
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, p.tmem_cols);
  if (clustered) cluster_sync_all();  // nobody leaves while a peer may still read its row statistics
  LIN_TS(17);
}

// W (N, K) row-major [* scale[n]] -> blocks [kb] of float4 [kcnt/4][N] (K-major core-matrix layout,
// chunk-major), rounded to tf32 once.  K is zero-padded to a multiple of 8.
__global__ void pack_linear_kernel(const float *__restrict__ W, const float *__restrict__ scale, int N, int K, int Kpad,
                                   int NS, long long sn, long long sk, float *__restrict__ dst) {
  // destination order: [column split s][k block][16-byte chunk][NS columns] float4
  const int total = N * Kpad;
  const int per_split = Kpad * NS;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int sidx = e / per_split;
    const int e1 = e - sidx * per_split;
    const int kb = e1 / (kKBlock * NS);
    const int rem = e1 - kb * kKBlock * NS;
    const int c = rem / (4 * NS);
    const int n = sidx * NS + ((rem >> 2) % NS);
    const int k = kb * kKBlock + c * 4 + (rem & 3);
    float w = 0.f;
    if (k < K) {
      w = W[n * sn + k * sk];
      if (scale) w *= scale[n];
    }
    dst[e] = to_tf32(w);
  }
}

// Batched variant: block (x, y) packs part of weight y of a descriptor array in device memory.  One launch re-packs
// every weight of a model (a captured training step does this once per replay instead of ~370 single launches).
struct PackDesc {
  const float *w;
  float *dst;
  long long sn, sk;
  int N, K, Kpad, pad_;
};
__global__ void pack_linear_batch_kernel(const PackDesc *__restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  const int total = d.N * d.Kpad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kb = e / (kKBlock * d.N);
    const int rem = e - kb * kKBlock * d.N;
    const int c = rem / (4 * d.N);
    const int n = (rem >> 2) % d.N;
    const int k = kb * kKBlock + c * 4 + (rem & 3);
    d.dst[e] = to_tf32(k < d.K ? d.w[n * d.sn + k * d.sk] : 0.f);
  }
}

// How many CTAs share the N columns of a row tile.  A 128 x 288 x 288 tile moves A 147 KB [+ pos], W 332 KB, residual and
// output 147 KB each through ONE SM, which sustains ~30 B/clk to and from L2 (scripts/lin_ts.py: K loop 14k cycles,
// residual pass 8k, store pass 5.5k of a 36k-cycle tile), so a launch with few row tiles is bound by per-SM bandwidth
// while most of the chip idles.  Splitting N over S CTAs divides the weight / residual / output traffic per SM by S (the
// A tile is re-staged by every slice).  It only pays while all tiles x S CTAs are resident at once, so S is chosen per
// launch from the tile count; the packed weight layout does not depend on it (a slice's columns of one 16-byte chunk row
// are contiguous: kcnt / 4 bulk copies per k block instead of one).
// EDA_LINEAR_SPLIT=0 disables splitting, =1 forces the widest split regardless of the tile count (tests).
inline int lin_splits(int N, int tiles, int sms) {
  static const int mode = [] { const char *e = getenv("EDA_LINEAR_SPLIT"); return e ? atoi(e) : -1; }();
  if (mode == 0) return 1;
  for (int S = 4; S >= 2; --S) {
    if (N % S || (N / S) % 16 || N / S < 48) continue;
    if (mode == 1 || (long long)tiles * S <= sms) return S;
  }
  return 1;
}

// keep-mask (1.0 / 0.0) of the dropout decisions above, for the backward pass: out[a * cols + b]
__global__ void dropout_mask_kernel(uint32_t seed_base, uint32_t thresh, long long rows, int cols, uint32_t a_mul, uint32_t a_add,
                                    float *__restrict__ out, const uint32_t *seed_epoch) {
  const uint32_t seed = effective_seed(seed_base, seed_epoch);
  const long long total = rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long a = e / cols;
    const int b = (int)(e - a * cols);
    out[e] = dropout_keep(seed, (uint32_t)a * a_mul + a_add, (uint32_t)b, thresh) ? 1.f : 0.f;
  }
}

// nn.Dropout as its own pass (after a BatchNorm + ReLU apply, where no GEMM epilogue is available to carry it): same
// counter-based decision as above, out = keep ? x / (1 - p) : 0
__global__ void dropout_apply_kernel(const float *__restrict__ x, uint32_t seed_base, uint32_t thresh, float scale,
                                     long long rows, int cols, uint32_t a_mul, uint32_t a_add, float *__restrict__ out,
                                     const uint32_t *seed_epoch) {
  const uint32_t seed = effective_seed(seed_base, seed_epoch);
  const long long total = rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long a = e / cols;
    const int b = (int)(e - a * cols);
    out[e] = dropout_keep(seed, (uint32_t)a * a_mul + a_add, (uint32_t)b, thresh) ? __ldg(x + e) * scale : 0.f;
  }
}

inline int kpad_of(int K) { return (K + 7) & ~7; }
// Tensor map over x (rows, K) fp32 row-major for 32-column x 128-row boxes, SWIZZLE_128B.  false = use cp.async staging.
inline bool make_a_tensor_map(CUtensorMap *map, const float *x, long long rows, int K) {
  if (K & 3) return false;
  return make_tensor_map_rows32(map, x, rows, K, K, kRows);
}

inline bool lin_supported(int N, int K) { return N >= 16 && N <= kMaxN && (N & 15) == 0 && K >= 1 && K <= 4096; }

}  // namespace
}  // namespace eda

extern "C" {

int eda_debug_timestamps(long long *host_out, int n) {
  if (n < 0) return eda::wgrad_tc_timestamps(host_out, -n);  // negative count: the weight-gradient kernel's phase stamps
  if (!host_out || n > 128) return EDA_ERR_INVALID_ARGUMENT;
  EDA_CUDA_TRY(cudaMemcpyFromSymbol(host_out, eda::g_lin_ts, sizeof(long long) * n), "debug timestamps");
  return EDA_OK;
}

int eda_dropout_mask(unsigned int seed, const unsigned int *dropout_epoch, float p, long long rows, int cols,
                     unsigned int a_mul, unsigned int a_add, float *out, void *stream) {
  using namespace eda;
  if (rows < 0 || cols < 0 || p < 0.f || p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0 || cols == 0) return EDA_OK;
  if (!out) return EDA_ERR_INVALID_ARGUMENT;
  const long long total = rows * cols;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  dropout_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(seed, dropout_thresh(p), rows, cols, a_mul, a_add, out,
                                                           reinterpret_cast<const uint32_t *>(dropout_epoch));
  return check_launch("dropout_mask_kernel");
}

int eda_dropout_apply(const float *x, unsigned int seed, const unsigned int *dropout_epoch, float p, long long rows, int cols,
                      unsigned int a_mul, unsigned int a_add, float *out, void *stream) {
  using namespace eda;
  if (rows < 0 || cols < 0 || p < 0.f || p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0 || cols == 0) return EDA_OK;
  if (!x || !out) return EDA_ERR_INVALID_ARGUMENT;
  const long long total = rows * cols;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  dropout_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, seed, dropout_thresh(p), 1.0f / (1.0f - p), rows, cols, a_mul,
                                                            a_add, out, reinterpret_cast<const uint32_t *>(dropout_epoch));
  return check_launch("dropout_apply_kernel");
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
  pack_linear_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(W, scale, N, K, kpad_of(K), N,
                                                                         (long long)K, 1LL, packed);
  return check_launch("pack_linear_kernel");
}

int eda_linear_pack_strided(const float *W, long long stride_n, long long stride_k, int N, int K, float *packed,
                            void *stream) {
  using namespace eda;
  if (!lin_supported(N, K)) return EDA_ERR_UNSUPPORTED;
  if (!W || !packed) return EDA_ERR_INVALID_ARGUMENT;
  const int total = N * kpad_of(K);
  pack_linear_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(W, nullptr, N, K, kpad_of(K), N,
                                                                         stride_n, stride_k, packed);
  return check_launch("pack_linear_kernel");
}

int eda_linear_pack_batch(const void *descs_device, int count, int max_elements, void *stream) {
  using namespace eda;
  static_assert(sizeof(PackDesc) == sizeof(eda_linear_pack_desc), "descriptor layout");
  if (count < 0 || max_elements < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (count == 0 || max_elements == 0) return EDA_OK;
  if (!descs_device || count > 65535) return EDA_ERR_INVALID_ARGUMENT;
  int bx = (max_elements + 255) / 256;
  if (bx > 64) bx = 64;
  pack_linear_batch_kernel<<<dim3((unsigned)bx, (unsigned)count), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const PackDesc *>(descs_device));
  return check_launch("pack_linear_batch_kernel");
}

int eda_linear_forward(const eda_linear_problem *probs, int nprobs, int K, int N, int relu, const float *ln_gamma,
                       const float *ln_beta, float ln_eps, int layer_norm, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch,
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
    if (q.pre_ln && (q.y_batch_rows > 0 || (reinterpret_cast<uintptr_t>(q.pre_ln) & 15))) return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].x = q.x; p.pr[i].pos = q.pos; p.pr[i].w = q.w_packed; p.pr[i].bias = q.bias;
    p.pr[i].residual = q.residual; p.pr[i].y = q.y; p.pr[i].pre = q.pre_ln; p.pr[i].rows = q.rows; p.pr[i].tile0 = tiles;
    p.pr[i].tb = q.y_batch_rows; p.pr[i].ldt = q.y_ld; p.pr[i].round_out = q.round_tf32;
    p.pr[i].tma = (q.rows > 0 && (!q.pos || !(reinterpret_cast<uintptr_t>(q.pos) & 15)) &&
                   make_a_tensor_map(&p.tmap[i], q.x, q.rows, K)) ? 1 : 0;
    if (q.y_row_stride != 0 && (q.y_row_stride < N || (q.y_row_stride & 3) || q.y_batch_rows > 0)) return EDA_ERR_INVALID_ARGUMENT;
    p.pr[i].ldy = q.y_row_stride ? q.y_row_stride : N;
    tiles += (q.rows + kRows - 1) / kRows;
  }
  if (tiles == 0) return EDA_OK;
  p.nprobs = nprobs; p.K = K; p.Kpad = kpad_of(K); p.N = N; p.relu = relu; p.ln = layer_norm ? 1 : 0;
  p.gamma = ln_gamma; p.beta = ln_beta; p.eps = ln_eps;
  if (dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  p.drop_thresh = dropout_thresh(dropout_p); p.drop_seed = dropout_seed; p.drop_scale = 1.0f / (1.0f - dropout_p);
  p.seed_epoch = reinterpret_cast<const uint32_t *>(dropout_epoch);
  const int sms = sm_count();
  p.S = lin_splits(N, tiles, sms);
  p.NS = N / p.S;
  const int NS = p.NS;
  p.tmem_cols = NS <= 32 ? 32u : NS <= 64 ? 64u : NS <= 128 ? 128u : NS <= 256 ? 256u : 512u;
  // stage = A tile + W block, padded to 1024 bytes (SWIZZLE_128B atoms must stay 1024-aligned).  Four stages (the
  // staging warps then run two K blocks ahead of the MMAs) whenever they fit next to the ~5 KB of static shared memory:
  // N = 288 -> 4 x 52 KB = 208 KB
  // Wide stages (sub > 1) need every problem's A tiles to come by TMA; the cp.async / unaligned staging paths keep the
  // 32-wide stages of before.  EDA_LINEAR_SUB forces a width (tests / measurements).
  bool all_tma = true;
  for (int i = 0; i < nprobs; ++i) all_tma = all_tma && (p.pr[i].rows == 0 || p.pr[i].tma);
  static const int sub_env = [] { const char *e = getenv("EDA_LINEAR_SUB"); return e ? atoi(e) : 0; }();
  p.sub = 1;
  if (all_tma) {
    const int nkb32 = (p.Kpad + kKBlock - 1) / kKBlock;
    for (int sub = 3; sub >= 2; --sub) {
      if (sub_env && sub != sub_env) continue;
      const size_t sb = ((size_t)sub * (kTileBytes + kKBlock * NS * 4) + 1023) & ~(size_t)1023;
      if (2 * sb <= 216 * 1024 && nkb32 >= 2) { p.sub = sub; break; }
    }
    if (sub_env == 1) p.sub = 1;
  }
  p.a_bytes = (uint32_t)(p.sub * kTileBytes);
  p.stage_bytes = (uint32_t)((p.a_bytes + p.sub * kKBlock * NS * 4 + 1023) & ~1023);
  if (p.sub > 1) {
    p.nstages = (size_t)3 * p.stage_bytes <= 216 * 1024 ? 3 : 2;
    p.prefetch = 0;  // TMA-fed: the producer runs ahead as far as the ring allows
  } else {
    p.nstages = (size_t)4 * p.stage_bytes <= 216 * 1024 ? 4 : 3;
    p.prefetch = p.nstages - 2;
  }
  {
    const char *e = getenv("EDA_LIN_TSHACK");
    p.ts_hack = e ? atoi(e) : 0;
  }
  size_t smem = (size_t)p.nstages * p.stage_bytes;
  const size_t out_tile = (size_t)kRows * (NS + 4) * sizeof(float);
  if (smem < out_tile) smem = out_tile;
  smem += 1024;  // alignment slack
  static SmemAttr attr;  // raised once per size increase and device
  EDA_CUDA_TRY(attr.ensure(linear_kernel, smem), "linear smem attr");
  if (p.ln && p.S > 1) {
    // LayerNorm needs whole rows: the S column slices of a row tile form a thread-block cluster and exchange
    // their per-row statistics through distributed shared memory
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * p.S));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    EDA_CUDA_TRY(cudaLaunchKernelEx(&cfg, linear_kernel, p), "linear_kernel cluster launch");
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * p.S));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    EDA_CUDA_TRY(cudaLaunchKernelEx(&cfg, linear_kernel, p), "linear_kernel launch");
  }
  return check_launch("linear_kernel");
}

}  // extern "C"
