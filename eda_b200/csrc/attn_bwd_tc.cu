// Backward of the attention core on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Measured (profiles/r1_ncu_attention_backward_tc.json, Nq = Nk = 1024, B = 8): 158 + 219 us, tcgen05 pipe 15 % active —
// with one CTA per SM the phases of a column block (staging wait, 10 MMAs, element-wise over 2 x 64 values per thread,
// 16-32 MMAs) run back to back, so the warp-level kernel of csrc/attn_bwd.cu (74 + 90 us, three CTAs per SM) is the
// default and this one is selected with EDA_ATTN_BWD=tc.  Next step: 64-column blocks (176 / 224 TMEM columns) so two
// CTAs per SM overlap each other's element-wise and MMA phases, as the forward kernel does.
//
// Same maths as csrc/attn_bwd.cu:
//     P  = exp(scale q k^T + mask - lse)      dP = dctx v^T       dS = P o (dP' - delta)
//     dq = scale dS k      dk = scale dS^T q      dv = P'^T dctx             (' = dropout keep / (1 - p) re-applied)
// as TWO launches of one kernel template, built from the blocks the forward kernel (csrc/attention.cu) established
// on the hardware: 128 rows = TMEM lanes, column blocks of 128, K-major SWIZZLE_NONE core-matrix operands staged by
// 16-byte cp.async one block ahead, one thread issuing tcgen05.mma kind::tf32, 256 threads sharing each row between
// two column halves for the element-wise stage.
//
//   rows = queries (dq):   S  = Q K^T   and  dP  = dO V^T   (A: shared-memory row tiles, B: natural K / V blocks)
//                          dS written IN PLACE over S in tensor memory, then  dq += dS K   (A: TMEM, B: channel-major K)
//   rows = keys (dk, dv):  S^T = K Q^T  and  dP^T = V dO^T  (the transposed orientation is simply recomputed)
//                          P' over S^T, dS^T over dP^T in place, then  dv += P'^T dO,  dk += dS^T Q
//                          (A: TMEM, B: channel-major dO / Q)
// No atomics, nothing of size Nq x Nk in HBM.  The B operand of the second products needs 4 consecutive column
// indices of one channel per 16-byte unit, i.e. a channel-major copy (B, H*D, ld) of k (first launch) and of q, dctx
// (second launch), prepared by the caller; rows beyond N inside a 4-group are zero-filled by the staging, padding
// columns are never read.  TMEM: 128 (S) + 128 (dP) + 64 + 64 (outputs) columns -> 512 allocated, one CTA per SM.
#include <math.h>
#include "umma.cuh"

namespace eda {
namespace {

constexpr int kRows = 128, kCB = 128, kThreads = 256;

struct AttnBwdTcParams {
  const float *q, *k, *v, *dctx, *ctx, *lse;   // natural (B, N, H*D); v scenes v_batch_stride floats apart
  const float *kt, *qt, *dot;                   // channel-major (B, H*D, ld): k (ldk), q / dctx (ldq)
  const unsigned char *mask;
  float *delta, *dq, *dk, *dv;
  long long v_batch_stride;
  int Nq, Nk, H, ldk, ldq;
  float scale;
  uint32_t drop_thresh, drop_seed;
  const uint32_t *seed_epoch;  // optional device word added to drop_seed (eda_dropout_set_epoch)
  float drop_scale;
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32x4(float4 a) {
  return make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct BDims {
  static constexpr int DC = D / 4;                // 16-byte chunks of real data per row
  static constexpr int QC = ((D + 7) / 8) * 2;    // chunks of the contraction depth of S / dP (multiple of 8 floats)
  static constexpr int VC = ((D + 15) / 16) * 4;  // chunks of the output width (multiple of 16 floats)
  static constexpr int Dn = VC * 4;
  static constexpr int VP = Dn + 1;               // float4 pitch of one 4-column group of a channel-major block
  static constexpr int VG = kCB / 4;
  static constexpr size_t smem_bytes(bool key_rows) {
    return (size_t)(2 * QC * kRows + 2 * 2 * QC * kCB + (key_rows ? 2 : 1) * 2 * VG * VP) * sizeof(float4) +
           2 * 2 * kCB * sizeof(float);
  }
};

template <int D, bool kKeyRows, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
attention_backward_tc_kernel(const AttnBwdTcParams p) {
  using DM = BDims<D>;
  constexpr int DC = DM::DC, QC = DM::QC, Dn = DM::Dn, VP = DM::VP, VG = DM::VG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *sX1 = reinterpret_cast<float4 *>(smem_raw);  // [QC][128] row tile of q (or k)
  float4 *sX2 = sX1 + QC * kRows;                      // [QC][128] row tile of dctx (or v)
  float4 *sC1 = sX2 + QC * kRows;                      // [2][QC][128] natural column block of k (or q)
  float4 *sC2 = sC1 + 2 * QC * kCB;                    // [2][QC][128] natural column block of v (or dctx)
  float4 *sT1 = sC2 + 2 * QC * kCB;                    // [2][VG][VP] channel-major column block of k (or q)
  float4 *sT2 = sT1 + 2 * VG * VP;                     // [2][VG][VP] channel-major dctx (rows = keys only)
  float *sStat = reinterpret_cast<float *>(kKeyRows ? sT2 + 2 * VG * VP : sT2);  // [2][2][128]
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_delta[kRows];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();  // == warp, known warp-uniform to the compiler (MMA issue branches)
  const int rt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int HD = p.H * D;
  const int Nrows = kKeyRows ? p.Nk : p.Nq, Ncols = kKeyRows ? p.Nq : p.Nk;
  const int nblocks = (Ncols + kCB - 1) / kCB;
  const size_t stat_base = ((size_t)b * p.H + h) * p.Nq;
  // natural tensors, this scene and head
  const float *qb = p.q + (size_t)b * p.Nq * HD + h * D, *kb = p.k + (size_t)b * p.Nk * HD + h * D;
  const float *vb = p.v + (size_t)b * p.v_batch_stride + h * D, *dob = p.dctx + (size_t)b * p.Nq * HD + h * D;
  const float *x1 = kKeyRows ? kb : qb, *x2 = kKeyRows ? vb : dob;
  const float *c1 = kKeyRows ? qb : kb, *c2 = kKeyRows ? dob : vb;
  const int ldt = kKeyRows ? p.ldq : p.ldk;
  const float *t1 = (kKeyRows ? p.qt : p.kt) + ((size_t)b * HD + h * D) * ldt;
  const float *t2 = kKeyRows ? p.dot + ((size_t)b * HD + h * D) * ldt : nullptr;
  constexpr float kLog2e = 1.4426950408889634f;

  const bool lo_side = tid < kRows;
  const int t128 = tid & (kRows - 1);
  const int tg = t128 & 31, tn0 = t128 >> 5;

  // column block `blk` -> buffer blk & 1: threads 0-127 natural C1 + channel-major T1, threads 128-255 C2 (+ T2)
  auto issue_cols = [&](int blk) {
    const int buf = blk & 1;
    const int col = blk * kCB + t128;
    const bool in = col < Ncols;
    {
      const float *src = (lo_side ? c1 : c2) + (size_t)(in ? col : 0) * HD;
      float4 *dst = (lo_side ? sC1 : sC2) + buf * QC * kCB + t128;
#pragma unroll
      for (int c = 0; c < DC; ++c) umma::cp_async16(dst + c * kCB, src + c * 4, in ? 16u : 0u);
    }
    if (lo_side || kKeyRows) {
      const float *tsrc = lo_side ? t1 : t2;
      float4 *tdst = (lo_side ? sT1 : sT2) + (buf * VG + tg) * VP;
      const int c0 = blk * kCB + 4 * tg;
      const int left = Ncols - c0;
      const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const int n = c * 4 + tn0;
        umma::cp_async16(tdst + n, bytes ? tsrc + (size_t)n * ldt + c0 : tsrc, bytes);
      }
    }
    umma::cp_async_commit();
    // per-column scalars of this block
    if (lo_side) {
      float *st = sStat + buf * 2 * kCB;
      if (kKeyRows) {
        st[t128] = in ? __ldg(p.lse + stat_base + col) * kLog2e : INFINITY;  // +inf: P = 0 for columns past the end
        st[kCB + t128] = in ? __ldg(p.delta + stat_base + col) : 0.f;
      } else {
        bool keep = in;
        if (keep && p.mask) keep = p.mask[(size_t)b * p.Nk + col] == 0;
        st[t128] = keep ? 0.f : -INFINITY;
      }
    }
  };

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&mma_done, 1);
    mbar_fence_init_cluster();
  }
  issue_cols(0);
  // zero padding of the staged blocks (depth D.. of natural blocks, widths D.. of channel-major blocks): written once
#pragma unroll
  for (int buf = 0; buf < 2; ++buf) {
#pragma unroll
    for (int c = DC; c < QC; ++c)
      (lo_side ? sC1 : sC2)[(buf * QC + c) * kCB + t128] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lo_side || kKeyRows) {
      float4 *tb = lo_side ? sT1 : sT2;
      for (int i = t128; i < VG * (Dn - D); i += kRows)
        tb[(buf * VG + i / (Dn - D)) * VP + D + i % (Dn - D)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- row tiles: threads 0-127 stage X1 (q or k), threads 128-255 X2 (dctx or v), one row each ------------------------
  {
    const int row = rt * kRows + t128;
    const bool in = row < Nrows;
    const float *src = (lo_side ? x1 : x2) + (size_t)(in ? row : 0) * HD;
    float4 *dst = (lo_side ? sX1 : sX2) + t128;
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < QC; ++c) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in && c < DC) {
        t = __ldg(reinterpret_cast<const float4 *>(src + c * 4));
        if (!kKeyRows && !lo_side) {  // delta = rowsum(dctx o ctx)
          const float4 o = __ldg(reinterpret_cast<const float4 *>(p.ctx + (size_t)b * p.Nq * HD + h * D + (size_t)row * HD + c * 4));
          dsum = fmaf(t.x, o.x, fmaf(t.y, o.y, fmaf(t.z, o.z, fmaf(t.w, o.w, dsum))));
        }
      }
      dst[c * kRows] = tf32x4(t);
    }
    if (!kKeyRows && !lo_side) {
      s_delta[t128] = dsum;
      if (in) p.delta[stat_base + row] = dsum;
    }
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  const int quad = warp & 3, half = warp >> 2;
  const int r = quad * 32 + lane;
  const int row = rt * kRows + r;
  const bool rvalid = row < Nrows;
  const uint32_t tS = umma::tmem_addr(tbase, (uint32_t)(quad * 32), (uint32_t)(half * 64));
  const uint32_t tP = tS + 128u;
  constexpr uint32_t kO1 = 256u, kO2 = 320u;

  // per-row scalars
  float row_l2 = 0.f, row_delta = 0.f, row_mask = 0.f;
  if (kKeyRows) {
    bool keep = rvalid;
    if (keep && p.mask) keep = p.mask[(size_t)b * p.Nk + row] == 0;
    row_mask = keep ? 0.f : -INFINITY;
  } else {
    row_l2 = rvalid ? __ldg(p.lse + stat_base + row) * kLog2e : INFINITY;
    row_delta = s_delta[r];
  }
  const float sl2 = p.scale * kLog2e;
  const uint32_t dseed = kDrop ? effective_seed(p.drop_seed, p.seed_epoch) : 0u;
  uint32_t phase = 0;

  for (int blk = 0; blk < nblocks; ++blk) {
    const int buf = blk & 1;
    const int c0 = blk * kCB;
    const int nc = min(kCB, Ncols - c0);
    const int ncp = (nc + 15) & ~15;
    if (blk + 1 < nblocks) {
      issue_cols(blk + 1);  // the other buffer was last read by the MMAs of block blk-1, which have completed
      umma::cp_async_wait<1>();
    } else {
      umma::cp_async_wait<0>();
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    // ---- S = X1 C1^T -> columns [0,128),  dP = X2 C2^T -> columns [128,256) ---------------------------------------------
    // (issue code: the whole of warp 0, converged, warp-uniform values, four K steps per statement — umma::mma4_tf32_ss_w)
    if (warp_u == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, ncp);
      const uint64_t a1 = umma::smem_desc_kmajor_noswizzle(smem_u32(sX1), kRows * 16u, 128u);
      const uint64_t a2 = umma::smem_desc_kmajor_noswizzle(smem_u32(sX2), kRows * 16u, 128u);
      const uint64_t b1 = umma::smem_desc_kmajor_noswizzle(smem_u32(sC1 + buf * QC * kCB), kCB * 16u, 128u);
      const uint64_t b2 = umma::smem_desc_kmajor_noswizzle(smem_u32(sC2 + buf * QC * kCB), kCB * 16u, 128u);
      constexpr uint32_t a_step = (2u * kRows * 16u) >> 4, b_step = (2u * kCB * 16u) >> 4;
      constexpr int kSteps = QC / 2, kGroups = kSteps / 4;
#pragma unroll
      for (int g = 0; g < kGroups; ++g)
        umma::mma4_tf32_ss_w(tbase, umma::desc_lo(a1) + (uint32_t)(4 * g) * a_step, umma::desc_hi(a1), a_step,
                             umma::desc_lo(b1) + (uint32_t)(4 * g) * b_step, umma::desc_hi(b1), b_step, idesc, g > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 4 * kGroups; ks < kSteps; ++ks)
        umma::mma_tf32_ss_w(tbase, a1 + ks * a_step, b1 + ks * b_step, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int g = 0; g < kGroups; ++g)
        umma::mma4_tf32_ss_w(tbase + 128u, umma::desc_lo(a2) + (uint32_t)(4 * g) * a_step, umma::desc_hi(a2), a_step,
                             umma::desc_lo(b2) + (uint32_t)(4 * g) * b_step, umma::desc_hi(b2), b_step, idesc, g > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 4 * kGroups; ks < kSteps; ++ks)
        umma::mma_tf32_ss_w(tbase + 128u, a2 + ks * a_step, b2 + ks * b_step, idesc, ks > 0 ? 1u : 0u);
      umma::mma_commit_w(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
    // ---- element-wise: this thread's 64 columns of its row, 16 at a time -----------------------------------------------
    const float *st = sStat + buf * 2 * kCB + half * 64;
    const int ncol = max(0, min(64, ncp - half * 64));
    // all 2 x 64 values of this thread in flight at once (one wait): two warps per scheduler cannot hide a TMEM round
    // trip per 16-column chunk
    uint32_t s[4][16], d[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 16 < ncol) {
        umma::tmem_ld16(tS + (uint32_t)(i * 16), s[i]);
        umma::tmem_ld16(tP + (uint32_t)(i * 16), d[i]);
      }
    }
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 16 < ncol) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int cl = i * 16 + e;  // column within this thread's half
          float l2, dl, madd;
          if (kKeyRows) {
            l2 = st[cl];
            dl = st[kCB + cl];
            madd = row_mask;
          } else {
            l2 = row_l2;
            dl = row_delta;
            madd = st[cl];
          }
          float pe = ex2_approx(fmaf(__uint_as_float(s[i][e]), sl2, madd) - l2);
          float dpe = __uint_as_float(d[i][e]);
          float pd = pe;
          if (kDrop) {
            const int col = c0 + half * 64 + cl;
            const int qi = kKeyRows ? col : row, ki = kKeyRows ? row : col;
            const bool keep = dropout_keep(dseed, (uint32_t)(stat_base + qi), (uint32_t)ki, p.drop_thresh);
            dpe = keep ? dpe * p.drop_scale : 0.f;
            pd = keep ? pe * p.drop_scale : 0.f;
          }
          const float ds = pe * (dpe - dl);
          if (kKeyRows) {
            s[i][e] = __float_as_uint(to_tf32(pd));   // P' over S^T
            d[i][e] = __float_as_uint(to_tf32(ds));   // dS^T over dP^T
          } else {
            s[i][e] = __float_as_uint(to_tf32(ds));   // dS over S
          }
        }
        umma::tmem_st16(tS + (uint32_t)(i * 16), s[i]);
        if (kKeyRows) umma::tmem_st16(tP + (uint32_t)(i * 16), d[i]);
      }
    }
    umma::tmem_st_wait();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    // ---- second products: A from tensor memory, B = channel-major column blocks ----------------------------------------------
    if (warp_u == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, Dn);
      const uint64_t tb1 = umma::smem_desc_kmajor_noswizzle(smem_u32(sT1 + buf * VG * VP), VP * 16u, 128u);
      const uint64_t tb2 = umma::smem_desc_kmajor_noswizzle(smem_u32(sT2 + buf * VG * VP), VP * 16u, 128u);
      constexpr uint32_t b_step = (2u * VP * 16u) >> 4;
      const int nsteps = ncp / 8;
      // key rows:   dk += dS^T Q (A = columns [128,256)),  dv += P'^T dO (A = columns [0,128))
      // query rows: dq += dS K   (A = columns [0,128))
      const uint32_t a_first = kKeyRows ? tbase + 128u : tbase;
      int ks = 0;
      for (; ks + 4 <= nsteps; ks += 4) {
        const uint32_t acc = (blk > 0 || ks > 0) ? 1u : 0u;
        umma::mma4_tf32_ts_w(tbase + kO1, a_first + (uint32_t)ks * 8u, 8u, umma::desc_lo(tb1) + (uint32_t)ks * b_step,
                             umma::desc_hi(tb1), b_step, idesc, acc);
        if (kKeyRows)
          umma::mma4_tf32_ts_w(tbase + kO2, tbase + (uint32_t)ks * 8u, 8u, umma::desc_lo(tb2) + (uint32_t)ks * b_step,
                               umma::desc_hi(tb2), b_step, idesc, acc);
      }
      for (; ks < nsteps; ++ks) {
        const uint32_t acc = (blk > 0 || ks > 0) ? 1u : 0u;
        umma::mma_tf32_ts_w(tbase + kO1, a_first + (uint32_t)ks * 8u, tb1 + (uint64_t)((uint32_t)ks * b_step), idesc, acc);
        if (kKeyRows)
          umma::mma_tf32_ts_w(tbase + kO2, tbase + (uint32_t)ks * 8u, tb2 + (uint64_t)((uint32_t)ks * b_step), idesc, acc);
      }
      umma::mma_commit_w(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
  }

  // ---- outputs: 16-column chunks dealt alternately to the two halves ----------------------------------------------------
  {
    constexpr int kOChunks = Dn / 16;
    float *o1 = (kKeyRows ? p.dk + (size_t)b * p.Nk * HD : p.dq + (size_t)b * p.Nq * HD) + (size_t)row * HD + h * D;
    float *o2 = kKeyRows ? p.dv + (size_t)b * p.Nk * HD + (size_t)row * HD + h * D : nullptr;
    const uint32_t tO = umma::tmem_addr(tbase, (uint32_t)(quad * 32), 0);
#pragma unroll
    for (int i = 0; i < kOChunks; ++i) {
      if ((i & 1) == half) {
        uint32_t u[16], w[16];
        umma::tmem_ld16(tO + kO1 + (uint32_t)(i * 16), u);
        if (kKeyRows) umma::tmem_ld16(tO + kO2 + (uint32_t)(i * 16), w);
        umma::tmem_ld_wait();
        if (rvalid) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (i * 16 + q4 * 4 < D) {
              *reinterpret_cast<float4 *>(o1 + i * 16 + q4 * 4) =
                  make_float4(__uint_as_float(u[q4 * 4 + 0]) * p.scale, __uint_as_float(u[q4 * 4 + 1]) * p.scale,
                              __uint_as_float(u[q4 * 4 + 2]) * p.scale, __uint_as_float(u[q4 * 4 + 3]) * p.scale);
              if (kKeyRows)
                *reinterpret_cast<float4 *>(o2 + i * 16 + q4 * 4) =
                    make_float4(__uint_as_float(w[q4 * 4 + 0]), __uint_as_float(w[q4 * 4 + 1]),
                                __uint_as_float(w[q4 * 4 + 2]), __uint_as_float(w[q4 * 4 + 3]));
            }
          }
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

template <int D, bool kKeyRows>
int launch_one(const AttnBwdTcParams &p, int B, cudaStream_t st) {
  const size_t smem = BDims<D>::smem_bytes(kKeyRows);
  static SmemAttr attr_plain, attr_drop;
  if (p.drop_thresh)
    EDA_CUDA_TRY(attr_drop.ensure(attention_backward_tc_kernel<D, kKeyRows, true>, smem), "attn bwd smem attr");
  else
    EDA_CUDA_TRY(attr_plain.ensure(attention_backward_tc_kernel<D, kKeyRows, false>, smem), "attn bwd smem attr");
  const int n = kKeyRows ? p.Nk : p.Nq;
  dim3 grid((unsigned)((n + kRows - 1) / kRows), (unsigned)p.H, (unsigned)B);
  if (p.drop_thresh)
    attention_backward_tc_kernel<D, kKeyRows, true><<<grid, kThreads, smem, st>>>(p);
  else
    attention_backward_tc_kernel<D, kKeyRows, false><<<grid, kThreads, smem, st>>>(p);
  return check_launch("attention_backward_tc_kernel");
}

template <int D>
int launch_both(const AttnBwdTcParams &p, int B, cudaStream_t st) {
  const int rc = launch_one<D, false>(p, B, st);
  if (rc != EDA_OK) return rc;
  return launch_one<D, true>(p, B, st);
}

}  // namespace
}  // namespace eda

extern "C" int eda_attention_backward_tc(const float *q, const float *k, const float *v, long long v_batch_stride,
                                         const float *kt, int ldk, const float *qt, const float *dctx_t, int ldq,
                                         const float *dctx, const float *ctx, const float *lse,
                                         const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                         float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *delta,
                                         float *dq, float *dk, float *dv, void *stream) {
  using namespace eda;
  if (B < 0 || Nq < 0 || Nk < 1 || H < 1 || D < 1) return EDA_ERR_INVALID_ARGUMENT;
  if (H > 65535 || B > 65535) return EDA_ERR_UNSUPPORTED;
  if (B == 0 || Nq == 0) return EDA_OK;
  if (!q || !k || !v || !kt || !qt || !dctx_t || !dctx || !ctx || !lse || !delta || !dq || !dk || !dv)
    return EDA_ERR_INVALID_ARGUMENT;
  if (ldk < Nk || (ldk & 3) || ldq < Nq || (ldq & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (v_batch_stride < (long long)Nk * H * D || (v_batch_stride & 3)) return EDA_ERR_INVALID_ARGUMENT;
  const void *ptrs[] = {q, k, v, kt, qt, dctx_t, dctx, ctx, dq, dk, dv};
  for (const void *ptr : ptrs)
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return EDA_ERR_INVALID_ARGUMENT;
  if (dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  AttnBwdTcParams p = {};
  p.q = q; p.k = k; p.v = v; p.dctx = dctx; p.ctx = ctx; p.lse = lse; p.kt = kt; p.qt = qt; p.dot = dctx_t;
  p.mask = key_padding_mask; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv; p.v_batch_stride = v_batch_stride;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.ldk = ldk; p.ldq = ldq; p.scale = scale;
  p.drop_thresh = dropout_thresh(dropout_p); p.drop_seed = dropout_seed; p.drop_scale = 1.0f / (1.0f - dropout_p);
  p.seed_epoch = reinterpret_cast<const uint32_t *>(dropout_epoch);
  cudaStream_t st = as_stream(stream);
  switch (D) {
    case 32: return launch_both<32>(p, B, st);
    case 36: return launch_both<36>(p, B, st);
    default: return EDA_ERR_UNSUPPORTED;
  }
}
