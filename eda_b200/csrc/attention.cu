// Multi-head attention core for sm_100a:  ctx = softmax(q k^T * scale + key_padding_mask) v
//
// Replaces the bmm / baddbmm -> softmax -> (dropout) -> bmm sequence of nn.MultiheadAttention's
// math path (torch/nn/functional.py:6607-6665; never the fused fast path in the reference because
// the modules are seq-first and need_weights is left True) for all 39 attention modules of
// models/encoder_decoder_layers.py (:87-93,99-105,111-117,149-153,179-183,366-370,374-380,385-391,
// 395-401).  The (B*H, Nq, Nk) score / probability tensors (vis self-attention: 2 x 268 MB at B=8)
// never exist in HBM and the head-averaged attention weights the reference computes and discards
// (need_weights=True) are not computed.
//
// CTA = (128-query tile, head, scene), 256 threads: query row = TMEM lane, and the two warps that share a
// lane quadrant split every 128-key block's columns between them.  Keys are processed in blocks of 128 with
// an online softmax:
//   K / V block  : 16-byte cp.async copies global -> shared, one block AHEAD of the maths (double
//                  buffered), straight into the tcgen05 K-major core-matrix layouts.  K (B,Nk,H*D) gives
//                  float4 [dim/4][key]; V arrives CHANNEL-major (B,H*D,ldv) — written that way by the
//                  V-projection GEMM's epilogue — so four consecutive keys of one channel are already one
//                  16-byte unit of the PV "B" operand and no transpose happens anywhere.  Both were rounded
//                  to tf32 (round-to-nearest) by that epilogue, so the tiles are used as they land
//   S[128 x nk]  = Q_h[128 x 40] K_h[nk x 40]^T     tcgen05.mma kind::tf32, head dim 36 zero-padded to
//                                                   40, q pre-multiplied by `scale`
//   row softmax  : each thread pulls its 64 scores into registers (one TMEM read pass), the two halves of a
//                  row exchange their maxima through shared memory, running max m and sum l stay in
//                  registers, P = exp(S + mask - m) is written back IN PLACE (tf32) over S
//   O[128 x 48] += P[128 x nk] V_h[nk x 48]         A operand = P straight from tensor memory;
//                                                   O rescaled by exp(m_old - m_new) between blocks
//   ctx          = O / l                            written to (B, Nq, H*D), 16-byte stores
// TMEM: 128 columns S/P + 48 columns O -> 256 allocated, two CTAs per SM overlap each other's
// staging / softmax / MMA phases.
// A fully masked row gives NaN (0/0) like the reference's softmax over all -inf.
#include <math.h>
#include "umma.cuh"

namespace eda {
namespace {

constexpr int kRows = 128;  // queries per CTA
constexpr int kKB = 128;    // keys per block
constexpr int kThreads = 256;  // 8 warps: two per TMEM lane quadrant

struct AttnParams {
  const float *q, *k, *vt;    // vt: (B, H*D, ldv) channel-major
  const unsigned char *mask;  // (B, Nk), nonzero = key ignored; may be null
  float *ctx;
  float *lse;                 // optional (B, H, Nq): log-sum-exp of the masked scores, for eda_attention_backward
  int Nq, Nk, H, ldv;
  float scale;
  uint32_t drop_thresh, drop_seed;  // dropout on the attention probabilities (nn.MultiheadAttention(dropout=p)); 0 = off
  const uint32_t *seed_epoch;       // optional device word added to drop_seed (eda_dropout_set_epoch)
  float drop_scale;
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 2^x on the SFU (MUFU.EX2, <= 2 ulp; -inf -> 0): softmax numerators need nothing more
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 tf32x4(float4 a) {
  return make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
}

template <int D>
struct Dims {
  static constexpr int DC = D / 4;              // 16-byte chunks of real data per row
  static constexpr int QC = ((D + 7) / 8) * 2;  // chunks of the QK^T depth (multiple of 8 floats)
  static constexpr int VC = ((D + 15) / 16) * 4;  // chunks of the PV width (multiple of 16 floats)
  static constexpr int Dn = VC * 4;
  static constexpr int VP = Dn + 1;  // float4 pitch of one 4-key group of V^T (odd: conflict-free cp.async stores)
  static constexpr int VG = kKB / 4; // 4-key groups per block
  static constexpr size_t smem_bytes =
      (size_t)(QC * kRows + 2 * QC * kKB + 2 * VG * VP) * sizeof(float4) + 2 * kKB * sizeof(float);
};

// Development aid: clock64() stamps of CTA (0,0,0), thread 0, key block 1 (eda_debug_timestamps_attn).
__device__ long long g_attn_ts[32];
#define ATT_TS(i) do { if (blk == 1 && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_attn_ts[i] = clock64(); } while (0)

template <int D, bool kDrop>
__global__ void __launch_bounds__(kThreads, 2)
attention_kernel(const AttnParams p) {
  pdl_launch_dependents();  // a PDL-launched successor may be scheduled now (it waits for this grid's completion itself)
  pdl_wait();               // launched with the PDL attribute: q / k / v come from the GEMM just before
  using DM = Dims<D>;
  constexpr int DC = DM::DC, QC = DM::QC, Dn = DM::Dn, VP = DM::VP, VG = DM::VG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *sQ = reinterpret_cast<float4 *>(smem_raw);  // [QC][128]
  float4 *sK = sQ + QC * kRows;                       // [2][QC][128]
  float4 *sV = sK + 2 * QC * kKB;                     // [2][VG][VP]: (dim n, key) -> float4 (key/4)*VP + n, lane key%4
  float *sMask = reinterpret_cast<float *>(sV + 2 * VG * VP);  // [2][128] additive 0 / -inf
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_x[2][kRows];  // per-row exchange between the two column halves (max, then sum)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();  // == warp, known warp-uniform to the compiler (MMA issue branches)
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int HD = p.H * D;
  const int nblocks = (p.Nk + kKB - 1) / kKB;
  const float *kbase = p.k + (size_t)b * p.Nk * HD + h * D;
  const float *vbase = p.vt + ((size_t)b * HD + h * D) * p.ldv;
  // staging ownership: threads 0-127 copy K (thread = key, DC chunks), threads 128-255 copy V^T
  // (chunk id = i * 128 + t -> dim n = id / 32, 4-key group g = id % 32: a warp moves 512 contiguous bytes)
  const bool k_side = tid < kRows;
  const int t128 = tid & (kRows - 1);
  const int vg = t128 & 31, vn0 = t128 >> 5;

  auto issue_kv = [&](int blk) {
    const int buf = blk & 1;
    if (k_side) {
      const int key = blk * kKB + t128;
      const bool in = key < p.Nk;
      const float *ks = in ? kbase + (size_t)key * HD : kbase;
      float4 *dk = sK + buf * QC * kKB + t128;
#pragma unroll
      for (int c = 0; c < DC; ++c) umma::cp_async16(dk + c * kKB, ks + c * 4, in ? 16u : 0u);
    } else {
      const int vkey = blk * kKB + 4 * vg;
      const int left = p.Nk - vkey;  // keys of this 4-key group that exist
      const uint32_t vbytes = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
      float4 *dv = sV + (buf * VG + vg) * VP;
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const int n = c * 4 + vn0;
        umma::cp_async16(dv + n, vbytes ? vbase + (size_t)n * p.ldv + vkey : vbase, vbytes);
      }
    }
    umma::cp_async_commit();
  };

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    mbar_init(&mma_done, 1);
    mbar_fence_init_cluster();
  }
  issue_kv(0);
  // zero padding (dims D.. of the padded depth / width) in both buffers: written once, never overwritten
#pragma unroll
  for (int buf = 0; buf < 2; ++buf) {
    if (k_side) {
#pragma unroll
      for (int c = DC; c < QC; ++c) sK[(buf * QC + c) * kKB + t128] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = t128; i < VG * (Dn - D); i += kRows)
        sV[(buf * VG + i / (Dn - D)) * VP + D + i % (Dn - D)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- Q tile (scaled): threads 128-255, one row each -------------------------------------------------
  if (!k_side) {
    const int qr = qt * kRows + t128;
    const float *src = p.q + ((size_t)b * p.Nq + qr) * HD + h * D;
#pragma unroll
    for (int c = 0; c < QC; ++c) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qr < p.Nq && c < DC) {
        t = __ldg(reinterpret_cast<const float4 *>(src + c * 4));
        t.x *= p.scale; t.y *= p.scale; t.z *= p.scale; t.w *= p.scale;
      }
      sQ[c * kRows + t128] = tf32x4(t);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  // softmax ownership: warp w reads TMEM lanes 32 (w % 4).. (its hardware quadrant) = query rows, and the
  // column half w / 4 of every 128-key block: two warps per scheduler hide each other's TMEM latency
  const int quad = warp & 3, half = warp >> 2;
  const int r = quad * 32 + lane;
  const int qrow = qt * kRows + r;
  const bool qvalid = qrow < p.Nq;
  const uint32_t tS = umma::tmem_addr(tbase, (uint32_t)(quad * 32), (uint32_t)(half * 64));
  const uint32_t tO = umma::tmem_addr(tbase, (uint32_t)(quad * 32), 128);
  // O columns owned by this thread (16-column chunks dealt alternately to the two halves)
  constexpr int kOChunks = Dn / 16;

  const uint32_t dseed = kDrop ? effective_seed(p.drop_seed, p.seed_epoch) : 0u;
  float m = -INFINITY, l = 0.f;  // l: this thread's share of the row sum (same running max in both halves)
  uint32_t phase = 0;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int blk = 0; blk < nblocks; ++blk) {
    const int buf = blk & 1;
    const int k0 = blk * kKB;
    const int nk = min(kKB, p.Nk - k0);
    const int nkp = (nk + 15) & ~15;
    ATT_TS(0);
    // the other buffer was last read by the MMAs of block blk-1, which have completed (waited below)
    if (blk + 1 < nblocks) {
      issue_kv(blk + 1);
      umma::cp_async_wait<1>();
    } else {
      umma::cp_async_wait<0>();
    }
    ATT_TS(1);
    // ---- K and V^T arrive already rounded to tf32 by the projection GEMM's epilogue (round_tf32): no fix-up.
    // Only the additive mask of this block is staged here.
    if (k_side) {
      bool keep = t128 < nk;
      if (keep && p.mask) keep = p.mask[(size_t)b * p.Nk + k0 + t128] == 0;
      sMask[buf * kKB + t128] = keep ? 0.f : -INFINITY;
    }
    ATT_TS(2);
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    ATT_TS(3);
    // ---- S = Q K^T --------------------------------------------------------------------------------
    // (issue code: the whole of warp 0, converged, warp-uniform values — see umma::mma4_tf32_ss_w)
    if (warp_u == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, nkp);
      const uint64_t ad = umma::smem_desc_kmajor_noswizzle(smem_u32(sQ), kRows * 16u, 128u);
      const uint64_t bd = umma::smem_desc_kmajor_noswizzle(smem_u32(sK + buf * QC * kKB), kKB * 16u, 128u);
      constexpr uint32_t a_step = (2u * kRows * 16u) >> 4, b_step = (2u * kKB * 16u) >> 4;
      constexpr int kSteps = QC / 2, kGroups = kSteps / 4;
#pragma unroll
      for (int g = 0; g < kGroups; ++g)
        umma::mma4_tf32_ss_w(tbase, umma::desc_lo(ad) + (uint32_t)(4 * g) * a_step, umma::desc_hi(ad), a_step,
                             umma::desc_lo(bd) + (uint32_t)(4 * g) * b_step, umma::desc_hi(bd), b_step, idesc, g > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 4 * kGroups; ks < kSteps; ++ks)
        umma::mma_tf32_ss_w(tbase, ad + ks * a_step, bd + ks * b_step, idesc, ks > 0 ? 1u : 0u);
      umma::mma_commit_w(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
    ATT_TS(4);
    // ---- online softmax: this thread's 64 columns of its row, held in registers ---------------------------
    const float *mk = sMask + buf * kKB + half * 64;
    const int ncol = max(0, min(64, nkp - half * 64));  // multiple of 16, uniform per warp
    uint32_t s[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i * 16 < ncol) umma::tmem_ld16(tS + (uint32_t)(i * 16), s[i]);
    umma::tmem_ld_wait();
    ATT_TS(5);
    float bm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 16 < ncol) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float sv = __uint_as_float(s[i][e]) + mk[i * 16 + e];
          s[i][e] = __float_as_uint(sv);
          bm = fmaxf(bm, sv);
        }
      }
    }
    s_x[half][r] = bm;
    ATT_TS(6);
    __syncthreads();
    ATT_TS(7);
    const float m_new = fmaxf(m, fmaxf(s_x[0][r], s_x[1][r]));
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = exp2f((m - m_use) * kLog2e);  // m = -inf -> 0
    const float moff = m_use * kLog2e;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 16 < ncol) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float pe = ex2_approx(fmaf(__uint_as_float(s[i][e]), kLog2e, -moff));
          sum += pe;  // the softmax denominator is taken BEFORE dropout, as in F.multi_head_attention_forward
          if (kDrop) {
            const uint32_t ra = (uint32_t)((b * p.H + h) * p.Nq + qrow);
            pe = dropout_keep(dseed, ra, (uint32_t)(k0 + half * 64 + i * 16 + e), p.drop_thresh) ? pe * p.drop_scale : 0.f;
          }
          s[i][e] = __float_as_uint(to_tf32(pe));
        }
        umma::tmem_st16(tS + (uint32_t)(i * 16), s[i]);
      }
    }
    l = l * alpha + sum;
    m = m_new;
    ATT_TS(8);
    if (blk > 0) {
#pragma unroll
      for (int i = 0; i < kOChunks; ++i) {
        if ((i & 1) == half) {
          uint32_t u[16];
          umma::tmem_ld16(tO + (uint32_t)(i * 16), u);
          umma::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) u[e] = __float_as_uint(__uint_as_float(u[e]) * alpha);
          umma::tmem_st16(tO + (uint32_t)(i * 16), u);
        }
      }
    }
    umma::tmem_st_wait();
    ATT_TS(9);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    ATT_TS(10);
    // ---- O += P V  (B operand = V^T block, K-major: 4-key units VP*16 bytes apart, dims 16 bytes apart) -----
    if (warp_u == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, Dn);
      const uint64_t bd = umma::smem_desc_kmajor_noswizzle(smem_u32(sV + buf * VG * VP), VP * 16u, 128u);
      constexpr uint32_t b_step = (2u * VP * 16u) >> 4;
      const int nsteps = nkp / 8;
      int ks = 0;
      for (; ks + 4 <= nsteps; ks += 4)
        umma::mma4_tf32_ts_w(tbase + 128u, tbase + (uint32_t)ks * 8u, 8u, umma::desc_lo(bd) + (uint32_t)ks * b_step,
                             umma::desc_hi(bd), b_step, idesc, (blk > 0 || ks > 0) ? 1u : 0u);
      for (; ks < nsteps; ++ks)
        umma::mma_tf32_ts_w(tbase + 128u, tbase + (uint32_t)ks * 8u, bd + (uint64_t)((uint32_t)ks * b_step), idesc,
                            (blk > 0 || ks > 0) ? 1u : 0u);
      umma::mma_commit_w(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
    ATT_TS(11);
  }

  // ---- ctx = O / l --------------------------------------------------------------------------------------
  {
    s_x[half][r] = l;
    __syncthreads();
    const float ltot = s_x[0][r] + s_x[1][r];
    const float inv = 1.0f / ltot;
    if (p.lse && half == 0 && qvalid) p.lse[((size_t)b * p.H + h) * p.Nq + qrow] = m + logf(ltot);
    float *dst = p.ctx + ((size_t)b * p.Nq + qrow) * HD + h * D;
#pragma unroll
    for (int i = 0; i < kOChunks; ++i) {
      if ((i & 1) == half) {
        uint32_t u[16];
        umma::tmem_ld16(tO + (uint32_t)(i * 16), u);
        umma::tmem_ld_wait();
        if (qvalid) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (i * 16 + q4 * 4 < D) {
              float4 o;
              o.x = __uint_as_float(u[q4 * 4 + 0]) * inv; o.y = __uint_as_float(u[q4 * 4 + 1]) * inv;
              o.z = __uint_as_float(u[q4 * 4 + 2]) * inv; o.w = __uint_as_float(u[q4 * 4 + 3]) * inv;
              *reinterpret_cast<float4 *>(dst + i * 16 + q4 * 4) = o;
            }
          }
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

template <int D>
int launch_attention(const AttnParams &p, int B, cudaStream_t st) {
  size_t smem = Dims<D>::smem_bytes;
  if (smem < 80 * 1024) smem = 80 * 1024;  // at most two CTAs per SM: their 2 x 256 TMEM columns always fit
  static SmemAttr attr_plain, attr_drop;  // per template instance (D), per device inside
  if (p.drop_thresh)
    EDA_CUDA_TRY(attr_drop.ensure(attention_kernel<D, true>, smem), "attention smem attr");
  else
    EDA_CUDA_TRY(attr_plain.ensure(attention_kernel<D, false>, smem), "attention smem attr");
  dim3 grid((unsigned)((p.Nq + kRows - 1) / kRows), (unsigned)p.H, (unsigned)B);
  if (p.drop_thresh)
    EDA_CUDA_TRY(launch_pdl(attention_kernel<D, true>, grid, dim3(kThreads), smem, st, p), "attention_kernel launch");
  else
    EDA_CUDA_TRY(launch_pdl(attention_kernel<D, false>, grid, dim3(kThreads), smem, st, p), "attention_kernel launch");
  return check_launch("attention_kernel");
}

}  // namespace
}  // namespace eda

extern "C" int eda_debug_timestamps_attn(long long *host_out, int n) {
  if (!host_out || n < 0 || n > 32) return EDA_ERR_INVALID_ARGUMENT;
  EDA_CUDA_TRY(cudaMemcpyFromSymbol(host_out, eda::g_attn_ts, sizeof(long long) * n), "debug timestamps");
  return EDA_OK;
}

extern "C" int eda_attention_forward_lse(const float *q, const float *k, const float *v, int ldv,
                                         const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                         float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *ctx,
                                         float *lse, void *stream) {
  using namespace eda;
  if (B < 0 || Nq < 0 || Nk < 1 || H < 1 || D < 1) return EDA_ERR_INVALID_ARGUMENT;
  if (H > 65535 || B > 65535) return EDA_ERR_UNSUPPORTED;
  if (ldv < Nk || (ldv & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (B == 0 || Nq == 0) return EDA_OK;
  if (!q || !k || !v || !ctx) return EDA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(ctx) & 15))
    return EDA_ERR_INVALID_ARGUMENT;
  AttnParams p = {};
  p.q = q; p.k = k; p.vt = v; p.mask = key_padding_mask; p.ctx = ctx; p.lse = lse;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.ldv = ldv; p.scale = scale;
  if (dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  p.drop_thresh = dropout_thresh(dropout_p); p.drop_seed = dropout_seed; p.drop_scale = 1.0f / (1.0f - dropout_p);
  p.seed_epoch = reinterpret_cast<const uint32_t *>(dropout_epoch);
  cudaStream_t st = as_stream(stream);
  switch (D) {  // head dims the compiled template set covers (EDA: 288 / 8 = 36)
    case 32: return launch_attention<32>(p, B, st);
    case 36: return launch_attention<36>(p, B, st);
    case 64: return launch_attention<64>(p, B, st);
    default: return EDA_ERR_UNSUPPORTED;
  }
}

extern "C" int eda_attention_forward(const float *q, const float *k, const float *v, int ldv,
                                     const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                     float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *ctx,
                                     void *stream) {
  return eda_attention_forward_lse(q, k, v, ldv, key_padding_mask, B, Nq, Nk, H, D, scale, dropout_p, dropout_seed,
                                   dropout_epoch, ctx, nullptr, stream);
}
