// Fused neighbour gather + grouped MLP + max-pool for the PointNet++ set-abstraction layer (sm_100a).
//
// Replaces, in one kernel, what the reference does in ~16 launches with every intermediate in HBM
// (pointnet2/pointnet2_utils.py:344-359 QueryAndGroup: 2x group_points, centre subtraction, /radius,
// concat;  pointnet2/pytorch_utils.py:11-36 SharedMLP: 3 x [1x1 Conv2d -> BatchNorm2d -> ReLU];
// pointnet2/pointnet2_modules.py:254-267 max_pool2d over nsample):
//
//   rows      = grouped points (b, centre j, sample s), 128 per tile  -> TMEM lanes (MMA M = 128)
//   layer l   : D_l[128 x C_l] = A_{l-1}[128 x K] * W'_l[C_l x K]^T    tcgen05.mma kind::tf32, fp32 accum
//   A_0       : [features(C) | (xyz - centre)/radius (3) | 0-pad]  gathered straight into TENSOR MEMORY
//   A_l       : relu(D_l + shift_l), written back IN PLACE in tensor memory (tcgen05.ld -> regs -> tcgen05.st)
//               and consumed by the next layer as the TMEM A operand: activations never touch
//               shared memory or HBM
//   W'_l      : BN-scale-folded weights, pre-packed into the K-major core-matrix layout and streamed
//               through a shared-memory ring by 1-D TMA bulk copies (one producer warp)
//   output    : max over the S rows of a centre: butterfly over lanes + red.max.s32 into a zeroed
//               (B, M, C3) buffer — the integer max against 0 IS the final ReLU
//
// BatchNorm: eval mode folds running statistics on the host side (scale into W', shift here).  Train
// mode needs batch statistics of every layer's conv output over all B*M*S rows: `stats_layer = l`
// runs layers < l normally and layer l with unscaled weights, then accumulates per-channel sum and
// sum of squares (butterfly + atomicAdd) instead of continuing; eda_bn_finalize turns them into
// scale/shift (and updates running stats) and the next pass goes one layer deeper.  4 passes, no
// (B,C,M,S) tensor in HBM.
//
// TMEM map per CTA (256 columns, two CTAs per SM):  [0,128) D1 -> A1, later D3 halves;
//                                                   [128,256) A0 chunks, later D2 -> A2.
#include "umma.cuh"

namespace eda {
namespace {

constexpr int kRows = 128;            // rows per tile = MMA M
constexpr int kComputeThreads = 128;  // warps 0-3: one thread per row
constexpr int kThreads = 160;         // + warp 4: weight producer
constexpr int kRing = 5;              // weight ring slots
constexpr int kSlotBytes = 16384;     // [8 chunks][128 n] float4
constexpr int kKBlock = 32;           // k per weight block
constexpr int kMaxC = 256;
constexpr unsigned kFull = 0xffffffffu;

struct SaParams {
  const float *xyz;      // (B,N,3)
  const float *new_xyz;  // (B,M,3)
  const float *feat;     // point-major features: row (b,i) at feat + (b*N+i)*feat_stride, C floats; null if C == 0
  const int *idx;        // (B,Mtot,S)
  const int *centre_idx; // optional (B,Mtot): centre j of scene b = xyz[b][centre_idx[b*Mtot+j]] instead of new_xyz
  int Mtot, m0;          // this launch covers centres m0 .. m0+M-1 of the Mtot centres of every scene
  const float *packed;   // pre-packed weights (eda_sa_mlp_pack)
  const float *shift[3]; // per-layer additive term after the (scale-folded) conv; null = 0
  float *out;            // (B,M,C3) zero-initialised, stats_layer == 0
  double *stats;         // [2][C_l] zero-initialised (sum, sum of squares; fp64 so that var = E[z^2] - E[z]^2 over ~1e6 rows keeps its digits), stats_layer > 0
  long long total_rows;  // B*M*S
  int N, M, S, log2S, C, feat_stride;
  int K0pad, Cout[3];
  int normalize;
  float radius;
  int stats_layer;       // 0 = full forward; 1..3 = stop after that layer's conv and accumulate stats
  int ntiles;
};

__device__ __forceinline__ void named_bar_sync_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Block sequence of one layer: for each N-half (128 columns), for each 32-wide k block.
__host__ __device__ inline int n_halves(int cout) { return (cout + 127) / 128; }
__host__ __device__ inline int n_kblocks(int k) { return (k + kKBlock - 1) / kKBlock; }

// Halving butterfly over the 16 values of a column group within each aligned 16-lane segment: every
// step exchanges half of the remaining values with the partner lane, so 15 shuffles reduce 16 columns
// over 16 rows and leave lane l with the result of column (l & 15).  OP is max or add.
template <bool kMax>
__device__ __forceinline__ float comb(float a, float b) {
  return kMax ? fmaxf(a, b) : a + b;
}
template <bool kMax>
__device__ __forceinline__ void butterfly16(float (&v)[16], int lane) {
  // masks 8,4,2,1: 16 -> 8 -> 4 -> 2 -> 1 values; lane l ends with column (l & 15) in v[0]
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8];
      const float keep = up ? v[i + 8] : v[i];
      v[i] = comb<kMax>(keep, __shfl_xor_sync(kFull, send, 8));
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = comb<kMax>(keep, __shfl_xor_sync(kFull, send, 4));
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = comb<kMax>(keep, __shfl_xor_sync(kFull, send, 2));
    }
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = comb<kMax>(keep, __shfl_xor_sync(kFull, send, 1));
  }
}

__global__ void __launch_bounds__(kThreads, 2)
sa_mlp_kernel(const SaParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *ring = reinterpret_cast<float4 *>(smem_raw);                       // [kRing][kSlotBytes]
  float *s_shift = reinterpret_cast<float *>(smem_raw + kRing * kSlotBytes);  // [3][kMaxC]
  __shared__ __align__(8) uint64_t full_bar[kRing], empty_bar[kRing], mma_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[4][2][kMaxC];  // per compute warp: this CTA's partial (sum, sum of squares) of the statistics pass

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();  // == warp, known warp-uniform to the compiler (MMA issue branches)
  const int nlayers_run = p.stats_layer == 0 ? 3 : p.stats_layer;

  // ---- one-time setup -----------------------------------------------------------------------
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&mma_done, 1);
    mbar_fence_init_cluster();
  }
  for (int i = tid; i < 3 * kMaxC; i += kThreads) {
    const int l = i / kMaxC, c = i % kMaxC;
    s_shift[i] = (p.shift[l] != nullptr && c < p.Cout[l]) ? __ldg(p.shift[l] + c) : 0.f;
  }
  for (int i = tid; i < 4 * 2 * kMaxC; i += kThreads) (&s_stat[0][0][0])[i] = 0.f;
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;

  // per-layer K (input width) as seen by the MMA
  int Kin[3] = {p.K0pad, p.Cout[0], p.Cout[1]};

  if (warp == 4) {
    // ================= weight producer: streams the packed blocks in consumption order ==========
    if (lane == 0) {
      uint32_t n = 0;  // running block counter (ring position)
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.packed);
        for (int l = 0; l < nlayers_run; ++l) {
          const int halves = n_halves(p.Cout[l]);
          const int kblocks = n_kblocks(Kin[l]);
          for (int h = 0; h < halves; ++h) {
            const int nblk = min(128, p.Cout[l] - h * 128);
            for (int kb = 0; kb < kblocks; ++kb) {
              const int kcnt = min(kKBlock, Kin[l] - kb * kKBlock);
              const uint32_t bytes = (uint32_t)kcnt * (uint32_t)nblk * 4u;
              const uint32_t slot = n % kRing;
              mbar_wait(&empty_bar[slot], ((n / kRing) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&full_bar[slot], bytes);
              bulk_g2s(reinterpret_cast<unsigned char *>(ring) + (size_t)slot * kSlotBytes, src, bytes,
                       &full_bar[slot]);
              src += bytes;
              ++n;
            }
          }
        }
      }
    }
  } else {
    // ================= compute warps: thread = row =================================================
    uint32_t n = 0;           // ring position; only warp 0 consumes blocks, it persists across tiles
    uint32_t done_phase = 0;  // parity of the next mma_done completion (same sequence in every thread)
    const uint32_t lane_base = (uint32_t)(warp * 32);
    const uint32_t t_r1 = umma::tmem_addr(tbase, lane_base, 0);    // region 1: columns [0,128)
    const uint32_t t_r2 = umma::tmem_addr(tbase, lane_base, 128);  // region 2: columns [128,256)
    const long long MS = (long long)p.M * p.S;

    // Warp 0 (all lanes, converged; one elected lane issues — umma::mma4_tf32_ts_w): D[128 x nblk] at column d_col (+)=
    // A[128 x (k_hi-k_lo)] (TMEM columns a_col...) * W-blocks^T.  Accumulates unless it is the very first k-step.
    auto issue_blocks = [&](uint32_t d_col, uint32_t a_col, int k_lo, int k_hi, int nblk, bool commit_done) {
      umma::fence_after_thread_sync();
      const uint32_t idesc = umma::idesc_tf32(kRows, nblk);
      const uint32_t lbo = (uint32_t)nblk * 16u;
      const uint32_t b_step = (2u * lbo) >> 4;
      const uint64_t bd0 = umma::smem_desc_kmajor_noswizzle(smem_u32(ring), lbo, 128u);
      for (int k0 = k_lo; k0 < k_hi; k0 += kKBlock) {
        const int kcnt = min(kKBlock, k_hi - k0);
        const uint32_t slot = n % kRing;
        mbar_wait(&full_bar[slot], (n / kRing) & 1u);
        umma::fence_after_thread_sync();
        const uint32_t b_lo = umma::desc_lo(bd0) + slot * (uint32_t)(kSlotBytes >> 4);
        const uint32_t a_t = tbase + a_col + (uint32_t)(k0 - k_lo);
        int ks = 0;
        for (; ks + 4 <= kcnt / 8; ks += 4)
          umma::mma4_tf32_ts_w(tbase + d_col, a_t + (uint32_t)ks * 8u, 8u, b_lo + (uint32_t)ks * b_step, umma::desc_hi(bd0),
                               b_step, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
        for (; ks < kcnt / 8; ++ks)
          umma::mma_tf32_ts_w(tbase + d_col, a_t + (uint32_t)ks * 8u,
                              ((uint64_t)umma::desc_hi(bd0) << 32) | (b_lo + (uint32_t)ks * b_step), idesc,
                              (k0 > 0 || ks > 0) ? 1u : 0u);
        umma::mma_commit_w(&empty_bar[slot]);  // slot reusable once these MMAs have read it
        ++n;
      }
      if (commit_done) umma::mma_commit_w(&mma_done);
    };
    auto wait_mma_done = [&]() {
      mbar_wait(&mma_done, done_phase);
      done_phase ^= 1u;
      umma::fence_after_thread_sync();
      __syncwarp();
    };
    // all compute threads: TMEM accesses of this phase are done / visible before thread 0 issues MMAs
    auto phase_sync = [&]() {
      umma::fence_before_thread_sync();
      named_bar_sync_compute();
    };
    // D_l -> A_l in place: relu(z + shift), rounded to tf32
    auto relu_inplace = [&](uint32_t taddr, int ncols, const float *sh) {
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(taddr + (uint32_t)c0, u);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 s4 = *reinterpret_cast<const float4 *>(sh + c0 + q4 * 4);
          u[q4 * 4 + 0] = __float_as_uint(to_tf32(fmaxf(__uint_as_float(u[q4 * 4 + 0]) + s4.x, 0.f)));
          u[q4 * 4 + 1] = __float_as_uint(to_tf32(fmaxf(__uint_as_float(u[q4 * 4 + 1]) + s4.y, 0.f)));
          u[q4 * 4 + 2] = __float_as_uint(to_tf32(fmaxf(__uint_as_float(u[q4 * 4 + 2]) + s4.z, 0.f)));
          u[q4 * 4 + 3] = __float_as_uint(to_tf32(fmaxf(__uint_as_float(u[q4 * 4 + 3]) + s4.w, 0.f)));
        }
        umma::tmem_st16(taddr + (uint32_t)c0, u);
      }
      umma::tmem_st_wait();
    };
    // per-channel sum / sum of squares of the raw conv output over this tile's valid rows
    // (fp32 partials per warp in shared memory, single writer per slot, summed in a fixed order — a CTA sees a few
    // thousand rows — then ONE fp64 atomic per channel and CTA at the end: per-tile atomics on the same 2 C addresses
    // serialise in L2, and fp32 atomics in arbitrary order make the statistics differ from run to run by ~1e-7, which
    // tf32 rounding of the activations amplifies to ~1e-3 on single outputs)
    auto take_stats = [&](uint32_t taddr, int ncols, float *sum, float *sumsq, bool valid) {
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(taddr + (uint32_t)c0, u);
        umma::tmem_ld_wait();
        float s1[16], s2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float z = valid ? __uint_as_float(u[e]) : 0.f;
          s1[e] = z;
          s2[e] = z * z;
        }
        butterfly16<false>(s1, lane);
        butterfly16<false>(s2, lane);
        const float a = s1[0] + __shfl_xor_sync(kFull, s1[0], 16);
        const float q = s2[0] + __shfl_xor_sync(kFull, s2[0], 16);
        if (lane < 16) {
          sum[c0 + lane] += a;
          sumsq[c0 + lane] += q;
        }
      }
    };

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      // ---- this thread's row ---------------------------------------------------------------
      const long long row = (long long)tile * kRows + tid;
      const bool valid = row < p.total_rows;
      long long b = 0, j = 0;
      float gx = 0.f, gy = 0.f, gz = 0.f;
      const float *frow = nullptr;
      if (valid) {
        b = row / MS;
        const long long rem = row - b * MS;
        j = p.m0 + (rem >> p.log2S);
        const int pi = __ldg(p.idx + (b * p.Mtot + j) * p.S + (rem & (p.S - 1)));
        const float *q = p.centre_idx ? p.xyz + (b * p.N + __ldg(p.centre_idx + b * p.Mtot + j)) * 3
                                      : p.new_xyz + (b * p.Mtot + j) * 3;
        const float *x = p.xyz + (b * p.N + pi) * 3;
        gx = __fsub_rn(__ldg(x), __ldg(q));  // pointnet2_utils.py:350  grouped_xyz -= new_xyz
        gy = __fsub_rn(__ldg(x + 1), __ldg(q + 1));
        gz = __fsub_rn(__ldg(x + 2), __ldg(q + 2));
        if (p.normalize) {  // pointnet2_utils.py:351-352  grouped_xyz /= radius
          gx = __fdiv_rn(gx, p.radius);
          gy = __fdiv_rn(gy, p.radius);
          gz = __fdiv_rn(gz, p.radius);
        }
        if (p.C > 0) frow = p.feat + (b * p.N + pi) * (long long)p.feat_stride;
      }

      // ================= layer 1: A0 chunks of <= 128 columns -> region 2, D1 -> region 1 ========
      const int nchunks = (p.K0pad + 127) / 128;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c_lo = ch * 128, c_hi = min(p.K0pad, c_lo + 128);
        if (ch > 0) wait_mma_done();  // MMAs reading the previous chunk finished: region 2 is free again
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
          uint32_t v[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int c = c0 + q4 * 4;
            float f[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
              if (c + 4 <= p.C && (p.feat_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feat) & 15) == 0) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(frow + c));
                f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int cc = c + e;
                  float val = 0.f;
                  if (cc < p.C) val = __ldg(frow + cc);
                  else if (cc == p.C) val = gx;
                  else if (cc == p.C + 1) val = gy;
                  else if (cc == p.C + 2) val = gz;
                  f[e] = val;
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[q4 * 4 + e] = __float_as_uint(to_tf32(f[e]));
          }
          umma::tmem_st16(t_r2 + (uint32_t)(c0 - c_lo), v);
        }
        umma::tmem_st_wait();
        phase_sync();
        if (warp_u == 0) issue_blocks(0u, 128u, c_lo, c_hi, p.Cout[0], true);
      }
      wait_mma_done();

      if (p.stats_layer == 1) {
        take_stats(t_r1, p.Cout[0], s_stat[warp][0], s_stat[warp][1], valid);
      } else {
        // ================= layer 2: A1 = relu(D1 + shift1) in region 1, D2 -> region 2 ===========
        relu_inplace(t_r1, p.Cout[0], s_shift);
        phase_sync();
        if (warp_u == 0) issue_blocks(128u, 0u, 0, p.Cout[0], p.Cout[1], true);
        wait_mma_done();
        if (p.stats_layer == 2) {
          take_stats(t_r2, p.Cout[1], s_stat[warp][0], s_stat[warp][1], valid);
        } else {
          // ================= layer 3: A2 in region 2, D3 halves -> region 1 ======================
          relu_inplace(t_r2, p.Cout[1], s_shift + kMaxC);
          phase_sync();
          const int C3 = p.Cout[2];
          const int halves = n_halves(C3);
          for (int h = 0; h < halves; ++h) {
            const int nblk = min(128, C3 - h * 128);
            if (warp_u == 0) issue_blocks(0u, 128u, 0, p.Cout[1], nblk, true);
            wait_mma_done();
            if (p.stats_layer == 3) {
              take_stats(t_r1, nblk, s_stat[warp][0] + h * 128, s_stat[warp][1] + h * 128, valid);
            } else {
              // shift, max over the S rows of each centre, ReLU via integer max against the zeroed output.
              // S >= 16 and a power of two: a centre's rows are one aligned 16-lane segment, one
              // warp, or several whole warps.
              const float *sh3 = s_shift + 2 * kMaxC + h * 128;
              int *orow = reinterpret_cast<int *>(p.out) + (valid ? (b * p.Mtot + j) * (long long)C3 + h * 128 : 0);
              for (int c0 = 0; c0 < nblk; c0 += 16) {
                uint32_t u[16];
                umma::tmem_ld16(t_r1 + (uint32_t)c0, u);
                umma::tmem_ld_wait();
                float y[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float4 s4 = *reinterpret_cast<const float4 *>(sh3 + c0 + q4 * 4);
                  y[q4 * 4 + 0] = __uint_as_float(u[q4 * 4 + 0]) + s4.x;
                  y[q4 * 4 + 1] = __uint_as_float(u[q4 * 4 + 1]) + s4.y;
                  y[q4 * 4 + 2] = __uint_as_float(u[q4 * 4 + 2]) + s4.z;
                  y[q4 * 4 + 3] = __uint_as_float(u[q4 * 4 + 3]) + s4.w;
                }
                butterfly16<true>(y, lane);  // lane l now holds column (l & 15) of its 16-lane segment
                float m = y[0];
                if (p.log2S >= 5) m = fmaxf(m, __shfl_xor_sync(kFull, m, 16));
                if (valid && (p.log2S == 4 || lane < 16)) atomicMax(orow + c0 + (lane & 15), __float_as_int(m));
              }
            }
            if (h + 1 < halves) phase_sync();  // the next half overwrites region 1
          }
        }
      }
      // every TMEM read of this tile is done before the next tile's stores / MMAs overwrite it
      phase_sync();
    }
    if (p.stats_layer > 0) {
      named_bar_sync_compute();  // all shared-memory partials are in
      const int Cl = p.Cout[p.stats_layer - 1];
      for (int c = tid; c < Cl; c += kComputeThreads) {
        atomicAdd(p.stats + c, ((double)s_stat[0][0][c] + (double)s_stat[1][0][c]) + ((double)s_stat[2][0][c] + (double)s_stat[3][0][c]));
        atomicAdd(p.stats + Cl + c, ((double)s_stat[0][1][c] + (double)s_stat[1][1][c]) + ((double)s_stat[2][1][c] + (double)s_stat[3][1][c]));
      }
    }
  }

  // ---- teardown -------------------------------------------------------------------------------
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

// ---- weight packing ---------------------------------------------------------------------------
// W (Cout, Kreal) row-major -> blocks [half][kblock] of float4 [kcnt/4][nblk]; column order of
// layer 1 is permuted to [features | xyz | pad] (the reference concatenates [xyz | features],
// pointnet2_utils.py:357-359; the sum over k does not care).  scale (per output channel, may be
// null) is folded in; values are rounded to tf32 (round-to-nearest) once, here.
__global__ void pack_layer_kernel(const float *__restrict__ W, const float *__restrict__ scale, int Cout, int Kreal,
                                  int Kpad, int first_layer_C, float *__restrict__ dst) {
  const int total = Cout * Kpad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    // e enumerates the destination linearly
    int rem = e;
    int h = 0, n0 = 0, nblk = 0;
    for (h = 0;; ++h) {
      n0 = h * 128;
      nblk = min(128, Cout - n0);
      const int half_elems = nblk * Kpad;
      if (rem < half_elems) break;
      rem -= half_elems;
    }
    int kb = 0, k0 = 0, kcnt = 0;
    for (kb = 0;; ++kb) {
      k0 = kb * kKBlock;
      kcnt = min(kKBlock, Kpad - k0);
      const int blk_elems = kcnt * nblk;
      if (rem < blk_elems) break;
      rem -= blk_elems;
    }
    const int chunk = rem / (nblk * 4);
    const int n = (rem / 4) % nblk;
    const int kk = k0 + chunk * 4 + (rem & 3);
    int ksrc = kk;  // column of the reference weight
    if (first_layer_C >= 0) {
      if (kk < first_layer_C) ksrc = kk + 3;                // features follow xyz in the reference
      else if (kk < first_layer_C + 3) ksrc = kk - first_layer_C;  // xyz
      else ksrc = -1;
    } else if (kk >= Kreal) {
      ksrc = -1;
    }
    float w = 0.f;
    if (ksrc >= 0 && ksrc < Kreal) {
      w = W[(size_t)(n0 + n) * Kreal + ksrc];
      if (scale) w *= scale[n0 + n];
    }
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w));
    dst[e] = __uint_as_float(r);
  }
}

// ---- BatchNorm statistics -> scale/shift --------------------------------------------------------
// stats = [sum(C), sumsq(C)] over `count` values.  Train-mode BatchNorm2d (pytorch_utils.py:40-64,
// eps 1e-5, momentum 0.1 set at models/bdetr.py:341-345): normalise with the biased variance, update
// running_var with the unbiased one.  Also used with precomputed running stats (count <= 0): then
// stats is ignored and running_mean/var are read.
__global__ void bn_finalize_kernel(const double *__restrict__ stats, double count, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum,
                                   float *__restrict__ running_mean, float *__restrict__ running_var, int update,
                                   int C, float *__restrict__ scale, float *__restrict__ shift,
                                   float *__restrict__ save_mean, float *__restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (count > 0) {
    mean = stats[c] / count;
    var = stats[C + c] / count - mean * mean;
    if (var < 0) var = 0;
    if (update && running_mean && running_var) {
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0;
  const double bt = beta ? (double)beta[c] : 0.0;
  scale[c] = (float)(g * invstd);
  shift[c] = (float)(bt - mean * g * invstd);
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = (float)invstd;
}

// ---- (B, R, C) <-> (B, C, R) -----------------------------------------------------------------
// `in` rows may be wider than C (row stride ld, first column col0): a column slice of a row-major matrix
__global__ void transpose_kernel(const float *__restrict__ in, int col0, int ld, int R, int C, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float *src = in + (size_t)b * R * ld + col0;
  float *dst = out + (size_t)b * R * C;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = src[(size_t)r * ld + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

size_t packed_floats(int C, int C1, int C2, int C3, int *K0pad_out) {
  const int K0pad = ((C + 3 + 15) / 16) * 16;
  if (K0pad_out) *K0pad_out = K0pad;
  return (size_t)C1 * K0pad + (size_t)C2 * C1 + (size_t)C3 * C2;
}

bool dims_supported(int C, int C1, int C2, int C3) {
  return C >= 0 && C1 >= 16 && C1 <= 128 && C1 % 16 == 0 && C2 >= 16 && C2 <= 128 && C2 % 16 == 0 && C3 >= 16 &&
         C3 <= 256 && C3 % 16 == 0 && (C3 <= 128 || C3 % 128 == 0 || (C3 - 128) % 16 == 0);
}

}  // namespace
}  // namespace eda

extern "C" {

size_t eda_sa_mlp_packed_floats(int C, int C1, int C2, int C3) {
  if (!eda::dims_supported(C, C1, C2, C3)) return 0;
  return eda::packed_floats(C, C1, C2, C3, nullptr);
}

int eda_sa_mlp_pack(const float *W1, const float *W2, const float *W3, const float *scale1, const float *scale2,
                    const float *scale3, int C, int C1, int C2, int C3, int nlayers, float *packed, void *stream) {
  using namespace eda;
  if (!dims_supported(C, C1, C2, C3)) return EDA_ERR_UNSUPPORTED;
  if (!W1 || !packed || nlayers < 1 || nlayers > 3 || (nlayers >= 2 && !W2) || (nlayers >= 3 && !W3))
    return EDA_ERR_INVALID_ARGUMENT;
  int K0pad = 0;
  packed_floats(C, C1, C2, C3, &K0pad);
  cudaStream_t st = as_stream(stream);
  float *dst = packed;
  pack_layer_kernel<<<(C1 * K0pad + 255) / 256, 256, 0, st>>>(W1, scale1, C1, C + 3, K0pad, C, dst);
  dst += (size_t)C1 * K0pad;
  if (nlayers >= 2) pack_layer_kernel<<<(C2 * C1 + 255) / 256, 256, 0, st>>>(W2, scale2, C2, C1, C1, -1, dst);
  dst += (size_t)C2 * C1;
  if (nlayers >= 3) pack_layer_kernel<<<(C3 * C2 + 255) / 256, 256, 0, st>>>(W3, scale3, C3, C2, C2, -1, dst);
  return check_launch("pack_layer_kernel", nlayers);
}

static int sa_mlp_forward_impl(const float *xyz, const float *new_xyz, const int *centre_idx, const float *feat,
                               int feat_stride, const int *idx, const float *packed, const float *shift1,
                               const float *shift2, const float *shift3, int B, int N, int M, int Mtot, int m0, int S,
                               int C, int C1, int C2, int C3, float radius, int normalize_xyz, int stats_layer,
                               int zero_fill, float *out, double *stats, void *stream) {
  using namespace eda;
  if (B < 0 || N <= 0 || M < 0 || S <= 0 || m0 < 0 || m0 + M > Mtot) return EDA_ERR_INVALID_ARGUMENT;
  if (!dims_supported(C, C1, C2, C3) || stats_layer < 0 || stats_layer > 3) return EDA_ERR_UNSUPPORTED;
  const int log2S = ilog2_exact(S);
  if (log2S < 4) return EDA_ERR_UNSUPPORTED;  // S must be a power of two >= 16 (rows of a centre = aligned lane segments)
  const long long total = (long long)B * M * S;
  if (total == 0) return EDA_OK;
  if (!xyz || (!new_xyz && !centre_idx) || !idx || !packed || (C > 0 && !feat) || (stats_layer == 0 ? !out : !stats))
    return EDA_ERR_INVALID_ARGUMENT;
  if (C > 0 && feat_stride < C) return EDA_ERR_INVALID_ARGUMENT;
  const long long ntiles = (total + kRows - 1) / kRows;
  if (ntiles > 0x7fffffffLL) return EDA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);

  SaParams p = {};
  p.xyz = xyz; p.new_xyz = new_xyz; p.centre_idx = centre_idx; p.feat = feat; p.idx = idx; p.packed = packed;
  p.shift[0] = shift1; p.shift[1] = shift2; p.shift[2] = shift3;
  p.out = out; p.stats = stats; p.total_rows = total;
  p.N = N; p.M = M; p.Mtot = Mtot; p.m0 = m0; p.S = S; p.log2S = log2S; p.C = C; p.feat_stride = feat_stride;
  packed_floats(C, C1, C2, C3, &p.K0pad);
  p.Cout[0] = C1; p.Cout[1] = C2; p.Cout[2] = C3;
  p.normalize = normalize_xyz; p.radius = radius; p.stats_layer = stats_layer; p.ntiles = (int)ntiles;

  const int Cl = stats_layer ? p.Cout[stats_layer - 1] : 0;
  if (zero_fill) {
    if (stats_layer == 0)
      EDA_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)B * Mtot * C3 * sizeof(float), st), "sa out memset");
    else
      EDA_CUDA_TRY(cudaMemsetAsync(stats, 0, (size_t)2 * Cl * sizeof(double), st), "sa stats memset");
  }

  const size_t smem = (size_t)kRing * kSlotBytes + 3 * kMaxC * sizeof(float);
  static SmemAttr attr;
  EDA_CUDA_TRY(attr.ensure(sa_mlp_kernel, smem), "sa smem attr");
  const int sms = sm_count();
  const int grid = (int)(ntiles < 2LL * sms ? ntiles : 2LL * sms);
  sa_mlp_kernel<<<grid, kThreads, smem, st>>>(p);
  return check_launch("sa_mlp_kernel");
}

int eda_sa_mlp_forward(const float *xyz, const float *new_xyz, const float *feat, int feat_stride, const int *idx,
                       const float *packed, const float *shift1, const float *shift2, const float *shift3, int B,
                       int N, int M, int S, int C, int C1, int C2, int C3, float radius, int normalize_xyz,
                       int stats_layer, float *out, double *stats, void *stream) {
  if (!new_xyz && (long long)B * M * S > 0) return EDA_ERR_INVALID_ARGUMENT;
  return sa_mlp_forward_impl(xyz, new_xyz, nullptr, feat, feat_stride, idx, packed, shift1, shift2, shift3, B, N, M, M, 0,
                             S, C, C1, C2, C3, radius, normalize_xyz, stats_layer, 1, out, stats, stream);
}

int eda_sa_mlp_forward_range(const float *xyz, const int *centre_idx, const float *feat, int feat_stride,
                             const int *idx, const float *packed, const float *shift1, const float *shift2,
                             const float *shift3, int B, int N, int Mtot, int m0, int mc, int S, int C, int C1, int C2,
                             int C3, float radius, int normalize_xyz, float *out, void *stream) {
  if (!centre_idx && (long long)B * mc * S > 0) return EDA_ERR_INVALID_ARGUMENT;
  return sa_mlp_forward_impl(xyz, nullptr, centre_idx, feat, feat_stride, idx, packed, shift1, shift2, shift3, B, N, mc,
                             Mtot, m0, S, C, C1, C2, C3, radius, normalize_xyz, 0, 0, out, nullptr, stream);
}

int eda_bn_finalize(const double *stats, double count, const float *gamma, const float *beta, float eps,
                    float momentum, float *running_mean, float *running_var, int update_running, int C,
                    float *scale, float *shift, float *save_mean, float *save_invstd, void *stream) {
  using namespace eda;
  if (C < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (C == 0) return EDA_OK;
  if (!scale || !shift || (count > 0 && !stats) || (count <= 0 && (!running_mean || !running_var)))
    return EDA_ERR_INVALID_ARGUMENT;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(stats, count, gamma, beta, eps, momentum,
                                                                     running_mean, running_var, update_running, C,
                                                                     scale, shift, save_mean, save_invstd);
  return check_launch("bn_finalize_kernel");
}

int eda_transpose_last2(const float *in, int B, int R, int C, float *out, void *stream) {
  using namespace eda;
  if (B < 0 || R < 0 || C < 0) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || R == 0 || C == 0) return EDA_OK;
  if (!in || !out) return EDA_ERR_INVALID_ARGUMENT;
  if (B > 65535 || (R + 31) / 32 > 65535) return EDA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B);
  transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, 0, C, R, C, out);
  return check_launch("transpose_kernel");
}

int eda_transpose_strided(const float *in, int col0, int ld, int B, int R, int C, float *out, void *stream) {
  using namespace eda;
  if (B < 0 || R < 0 || C < 0 || col0 < 0 || ld < col0 + C) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || R == 0 || C == 0) return EDA_OK;
  if (!in || !out) return EDA_ERR_INVALID_ARGUMENT;
  if (B > 65535 || (R + 31) / 32 > 65535) return EDA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B);
  transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, col0, ld, R, C, out);
  return check_launch("transpose_kernel");
}

}  // extern "C"
