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
// CTA = (128-query tile, head, scene); thread = query row = TMEM lane.  Keys are processed in blocks
// of 128 with an online softmax:
//   S[128 x nk]  = Q_h[128 x 40] K_h[nk x 40]^T     tcgen05.mma kind::tf32, operands staged in shared
//                                                   memory (K-major core-matrix layout), head dim 36
//                                                   zero-padded to 40, q pre-multiplied by `scale`
//   row softmax  : tcgen05.ld 16 columns at a time, running max m and sum l in registers,
//                  P = exp(S + mask - m) written back IN PLACE (tf32) over S
//   O[128 x 48] += P[128 x nk] V_h[nk x 48]         A operand = P straight from tensor memory,
//                                                   B = V block staged transposed in shared memory;
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
constexpr unsigned kFull = 0xffffffffu;

struct AttnParams {
  const float *q, *k, *v;
  const unsigned char *mask;  // (B, Nk), nonzero = key ignored; may be null
  float *ctx;
  int Nq, Nk, H, D, Dk, Dn;   // Dk = D rounded up to 8 (QK^T depth), Dn = D rounded up to 16 (PV width)
  float scale;
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32x4(float4 a) {
  return make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
}

__global__ void __launch_bounds__(kRows, 2)
attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Dk = p.Dk, Dn = p.Dn, D = p.D;
  float4 *sQ = reinterpret_cast<float4 *>(smem_raw);      // [Dk/4][128]
  float4 *sK = sQ + (Dk / 4) * kRows;                     // [Dk/4][128]
  float *sVt = reinterpret_cast<float *>(sK + (Dk / 4) * kKB);  // [128/4][Dn] float4: (n, key) -> (key/4)*Dn*4 + n*4 + key%4
  float *sMask = sVt + (kKB / 4) * Dn * 4;                // [128] additive 0 / -inf
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int HD = p.H * D;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    mbar_init(&mma_done, 1);
    mbar_fence_init_cluster();
  }
  // ---- Q tile (scaled) ---------------------------------------------------------------------------
  const int qrow = qt * kRows + tid;
  const bool qvalid = qrow < p.Nq;
  {
    const float *src = p.q + ((size_t)b * p.Nq + qrow) * HD + h * D;
    for (int c = 0; c < Dk / 4; ++c) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qvalid && c * 4 < D) {
        t = __ldg(reinterpret_cast<const float4 *>(src + c * 4));
        t.x *= p.scale; t.y *= p.scale; t.z *= p.scale; t.w *= p.scale;
      }
      sQ[c * kRows + tid] = tf32x4(t);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t tS = umma::tmem_addr(tbase, (uint32_t)(warp * 32), 0);
  const uint32_t tO = umma::tmem_addr(tbase, (uint32_t)(warp * 32), 128);

  float m = -INFINITY, l = 0.f;
  uint32_t phase = 0;
  const int nblocks = (p.Nk + kKB - 1) / kKB;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int blk = 0; blk < nblocks; ++blk) {
    const int k0 = blk * kKB;
    const int nk = min(kKB, p.Nk - k0);
    const int nkp = (nk + 15) & ~15;
    // ---- stage K block, V block (transposed), mask -------------------------------------------------
    {
      const int key = k0 + tid;
      const bool kvalid = tid < nk;
      const float *ksrc = p.k + ((size_t)b * p.Nk + key) * HD + h * D;
      const float *vsrc = p.v + ((size_t)b * p.Nk + key) * HD + h * D;
      for (int c = 0; c < Dk / 4; ++c) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid && c * 4 < D) t = __ldg(reinterpret_cast<const float4 *>(ksrc + c * 4));
        sK[c * kKB + tid] = tf32x4(t);
      }
      // rotated order: the 8 key groups of a warp hit 8 different bank quads -> conflict-free stores
      const int g = tid >> 2, r = tid & 3;
      float *dst = sVt + (size_t)g * Dn * 4 + r;
      int n = g % Dn;
      for (int i = 0; i < Dn; ++i) {
        float val = 0.f;
        if (kvalid && n < D) val = to_tf32(__ldg(vsrc + n));
        dst[n * 4] = val;
        n = (n + 1 == Dn) ? 0 : n + 1;
      }
      bool keep = kvalid;
      if (keep && p.mask) keep = p.mask[(size_t)b * p.Nk + key] == 0;
      sMask[tid] = keep ? 0.f : -INFINITY;
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    // ---- S = Q K^T --------------------------------------------------------------------------------
    if (tid == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, nkp);
      const uint32_t qb = smem_u32(sQ), kb = smem_u32(sK);
      for (int ks = 0; ks < Dk / 8; ++ks) {
        const uint64_t adesc = umma::smem_desc_kmajor_noswizzle(qb + (uint32_t)ks * 2u * kRows * 16u, kRows * 16u, 128u);
        const uint64_t bdesc = umma::smem_desc_kmajor_noswizzle(kb + (uint32_t)ks * 2u * kKB * 16u, kKB * 16u, 128u);
        umma::mma_tf32_ss(tbase, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
      }
      umma::mma_commit(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
    // ---- online softmax on this thread's row ----------------------------------------------------------
    float bm = -INFINITY;
    for (int c0 = 0; c0 < nkp; c0 += 16) {
      uint32_t u[16];
      umma::tmem_ld16(tS + (uint32_t)c0, u);
      umma::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) bm = fmaxf(bm, __uint_as_float(u[e]) + sMask[c0 + e]);
    }
    const float m_new = fmaxf(m, bm);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = exp2f((m - m_use) * kLog2e);  // m = -inf -> 0
    float sum = 0.f;
    for (int c0 = 0; c0 < nkp; c0 += 16) {
      uint32_t u[16];
      umma::tmem_ld16(tS + (uint32_t)c0, u);
      umma::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float pe = exp2f((__uint_as_float(u[e]) + sMask[c0 + e] - m_use) * kLog2e);
        sum += pe;
        u[e] = __float_as_uint(to_tf32(pe));
      }
      umma::tmem_st16(tS + (uint32_t)c0, u);
    }
    l = l * alpha + sum;
    m = m_new;
    if (blk > 0) {
      for (int c0 = 0; c0 < Dn; c0 += 16) {
        uint32_t u[16];
        umma::tmem_ld16(tO + (uint32_t)c0, u);
        umma::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) u[e] = __float_as_uint(__uint_as_float(u[e]) * alpha);
        umma::tmem_st16(tO + (uint32_t)c0, u);
      }
    }
    umma::tmem_st_wait();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    // ---- O += P V -----------------------------------------------------------------------------------
    if (tid == 0) {
      const uint32_t idesc = umma::idesc_tf32(kRows, Dn);
      const uint32_t vb = smem_u32(sVt);
      const uint32_t lbo = (uint32_t)Dn * 16u;
      for (int ks = 0; ks < nkp / 8; ++ks) {
        const uint64_t bdesc = umma::smem_desc_kmajor_noswizzle(vb + (uint32_t)ks * 2u * lbo, lbo, 128u);
        umma::mma_tf32_ts(tbase + 128u, tbase + (uint32_t)ks * 8u, bdesc, idesc, (blk > 0 || ks > 0) ? 1u : 0u);
      }
      umma::mma_commit(&mma_done);
    }
    mbar_wait(&mma_done, phase);
    phase ^= 1u;
    umma::fence_after_thread_sync();
    __syncwarp();
  }

  // ---- ctx = O / l --------------------------------------------------------------------------------------
  {
    const float inv = 1.0f / l;
    float *dst = p.ctx + ((size_t)b * p.Nq + qrow) * HD + h * D;
    for (int c0 = 0; c0 < Dn; c0 += 16) {
      uint32_t u[16];
      umma::tmem_ld16(tO + (uint32_t)c0, u);
      umma::tmem_ld_wait();
      if (qvalid) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (c0 + q4 * 4 < D) {
            float4 o;
            o.x = __uint_as_float(u[q4 * 4 + 0]) * inv; o.y = __uint_as_float(u[q4 * 4 + 1]) * inv;
            o.z = __uint_as_float(u[q4 * 4 + 2]) * inv; o.w = __uint_as_float(u[q4 * 4 + 3]) * inv;
            *reinterpret_cast<float4 *>(dst + c0 + q4 * 4) = o;
          }
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

}  // namespace
}  // namespace eda

extern "C" int eda_attention_forward(const float *q, const float *k, const float *v,
                                     const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                     float scale, float *ctx, void *stream) {
  using namespace eda;
  if (B < 0 || Nq < 0 || Nk < 1 || H < 1 || D < 1) return EDA_ERR_INVALID_ARGUMENT;
  if ((D & 3) || D > 64 || H > 65535 || B > 65535) return EDA_ERR_UNSUPPORTED;
  if (B == 0 || Nq == 0) return EDA_OK;
  if (!q || !k || !v || !ctx) return EDA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(ctx) & 15))
    return EDA_ERR_INVALID_ARGUMENT;
  AttnParams p = {};
  p.q = q; p.k = k; p.v = v; p.mask = key_padding_mask; p.ctx = ctx;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.D = D; p.Dk = (D + 7) & ~7; p.Dn = (D + 15) & ~15; p.scale = scale;
  size_t smem = (size_t)(p.Dk / 4) * kRows * 16 * 2 + (size_t)(kKB / 4) * p.Dn * 16 + kKB * sizeof(float);
  if (smem < 80 * 1024) smem = 80 * 1024;  // at most two CTAs per SM: their 2 x 256 TMEM columns always fit
  EDA_CUDA_TRY(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "attention smem attr");
  dim3 grid((unsigned)((Nq + kRows - 1) / kRows), (unsigned)H, (unsigned)B);
  attention_kernel<<<grid, kRows, smem, as_stream(stream)>>>(p);
  return check_launch("attention_kernel");
}
