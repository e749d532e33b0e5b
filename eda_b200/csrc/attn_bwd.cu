// Backward of the multi-head attention core (eda_attention_forward) for sm_100a:
//
//     given  q (B,Nq,H*D), k (B,Nk,H*D), v (B,Nk,H*D) projected,  ctx = softmax(q k^T scale + mask) v  (forward),
//            lse (B,H,Nq) = log-sum-exp of the masked scores (written by the forward kernel),  dctx
//     gives  dq, dk, dv
//
// In the reference this is autograd through nn.MultiheadAttention's math path (torch/nn/functional.py:6607-6665:
// bmm / softmax / dropout / bmm backward, five library kernels and two (B*H,Nq,Nk) tensors per module).  Here the
// probabilities are recomputed tile by tile from lse (flash-attention style), nothing of size Nq x Nk touches HBM:
//
//     P  = exp(S - lse),  S = scale q k^T + mask          dP = dctx v^T   (dropout: keep/(1-p) re-applied from the hash)
//     dS = P o (dP - delta),  delta = rowsum(dctx o ctx)
//     dq = scale dS k          dk = scale dS^T q          dv = P_drop^T dctx
//
// Two launches of ONE kernel template: rows = queries (accumulates dq over key blocks; also emits delta) and
// rows = keys (accumulates dk, dv over query blocks, S^T = k q^T recomputed in the transposed orientation), so no
// atomics and a deterministic result.  CTA = 64 rows x (head, scene), 4 warps x 16 rows; column blocks of 64 are
// staged by 16-byte cp.async (double buffered).  The contractions run on warp-level mma.sync.m16n8k8 tf32 with
// fp32 accumulation: the score tile is consumed as the A operand of the second product straight from the
// accumulator registers (the k index of that product is permuted so that accumulator columns 2t, 2t+1 are
// fragment columns t, t+4 — the B fragments are read from shared memory in the same order).
#include <math.h>
#include "common.cuh"

namespace eda {
namespace {

constexpr int kBwRows = 64, kBwCols = 64, kBwThreads = 128;
constexpr unsigned kFullMask = 0xffffffffu;

struct AttnBwdParams {
  const float *q, *k, *v, *dctx, *ctx, *lse;
  const unsigned char *mask;  // (B, Nk) nonzero = ignored
  float *delta;               // (B, H, Nq): written by the rows = queries launch, read by the rows = keys launch
  float *dq, *dk, *dv;
  long long v_batch_stride;   // floats between scenes of v (rows may be padded)
  const float *vt;            // alternative to v: channel-major values (B, H*D, ldv) exactly as the forward kernel took them
  int ldv;
  int Nq, Nk, H;
  float scale;
  uint32_t drop_thresh, drop_seed;
  const uint32_t *seed_epoch;  // optional device word added to drop_seed (eda_dropout_set_epoch)
  float drop_scale;
};

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kKeyRows = false: rows = queries, columns = keys, output dq (and delta).
// kKeyRows = true : rows = keys, columns = queries, outputs dk, dv.
template <int D, bool kKeyRows, bool kDrop>
__global__ void __launch_bounds__(kBwThreads)
attention_backward_kernel(const AttnBwdParams p) {
  pdl_launch_dependents();  // a PDL-launched successor may be scheduled now (it waits for this grid's completion itself)
  pdl_wait();               // launched with the PDL attribute: dctx comes from the GEMM just before, delta from the first pass
  constexpr int DP = (D + 7) & ~7;   // padded depth
  constexpr int KS = DP / 8;         // k-steps over the depth / n-tiles of the outputs
  constexpr int DC = D / 4;          // 16-byte chunks of real data per row
  constexpr int PITCH = DP + 4;      // (12 g + t) % 32 and (24 t + g) % 32 are conflict-free for PITCH = 44
  static_assert(D % 4 == 0, "head dim must be a multiple of 4");
  __shared__ __align__(16) float sC1[2][kBwCols][PITCH];  // rows=queries: K block; rows=keys: Q block
  constexpr int PT = kBwCols + 8;                         // pitch of the transposed V block: (8 t + g) % 32 conflict-free
  constexpr int C2N = (kBwCols * PITCH > DP * PT) ? kBwCols * PITCH : DP * PT;
  __shared__ __align__(16) float sC2[2][C2N];  // rows=queries: V block ([key][PITCH], or [dim][PT] from channel-major V); rows=keys: dctx block
  __shared__ float sStat[2][2][kBwCols];                  // [buf][0: additive mask or lse, 1: delta][col]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int HD = p.H * D;
  const int Nrows = kKeyRows ? p.Nk : p.Nq, Ncols = kKeyRows ? p.Nq : p.Nk;
  const int row0 = blockIdx.x * kBwRows + warp * 16;
  const float *x1base = kKeyRows ? p.k + (size_t)b * p.Nk * HD : p.q + (size_t)b * p.Nq * HD;
  const bool v_t = p.vt != nullptr;  // values arrive channel-major
  const float *vtb = v_t ? p.vt + ((size_t)b * HD + h * D) * p.ldv : nullptr;
  const float *x2base = kKeyRows ? p.v + (size_t)b * p.v_batch_stride : p.dctx + (size_t)b * p.Nq * HD;
  const float *c1base = kKeyRows ? p.q + (size_t)b * p.Nq * HD : p.k + (size_t)b * p.Nk * HD;
  const float *c2base = kKeyRows ? p.dctx + (size_t)b * p.Nq * HD : p.v + (size_t)b * p.v_batch_stride;
  const size_t stat_base = ((size_t)b * p.H + h) * p.Nq;
  const int nblocks = (Ncols + kBwCols - 1) / kBwCols;
  constexpr float kLog2e = 1.4426950408889634f;

  auto issue = [&](int blk, int buf) {
    const int c0 = blk * kBwCols;
    const bool c2_t = !kKeyRows && v_t;  // V block staged [dim][key] from the channel-major source
    for (int id = tid; id < kBwCols * DC; id += kBwThreads) {
      const int r = id / DC, ch = id - r * DC;
      const bool in = c0 + r < Ncols;
      const size_t off = (size_t)(c0 + r) * HD + h * D + ch * 4;
      cp_async16_zfill(&sC1[buf][r][ch * 4], in ? c1base + off : c1base, in ? 16u : 0u);
      if (!c2_t) cp_async16_zfill(&sC2[buf][r * PITCH + ch * 4], in ? c2base + off : c2base, in ? 16u : 0u);
    }
    if (c2_t) {
      for (int id = tid; id < D * (kBwCols / 4); id += kBwThreads) {
        const int d = id / (kBwCols / 4), kg = id - d * (kBwCols / 4);
        const int key = c0 + 4 * kg;
        const int left = Ncols - key;
        const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
        cp_async16_zfill(&sC2[buf][d * PT + 4 * kg], bytes ? vtb + (size_t)d * p.ldv + key : vtb, bytes);
      }
    }
    cp_async_commit_group();
    if (tid < kBwCols) {
      const int c = c0 + tid;
      const bool in = c < Ncols;
      if (kKeyRows) {
        // columns = queries: lse (+inf for columns past the end: P = 0) and delta
        sStat[buf][0][tid] = in ? __ldg(p.lse + stat_base + c) * kLog2e : INFINITY;
        sStat[buf][1][tid] = in ? __ldg(p.delta + stat_base + c) : 0.f;
      } else {
        bool keep = in;
        if (keep && p.mask) keep = p.mask[(size_t)b * p.Nk + c] == 0;
        sStat[buf][0][tid] = keep ? 0.f : -INFINITY;
      }
    }
  };

  // zero the padded depth columns D..DP-1 (+ pitch padding) once: cp.async never touches them
  for (int i = tid; i < 2 * kBwCols * (PITCH - D); i += kBwThreads) {
    const int bufr = i / (PITCH - D), c = D + i % (PITCH - D);
    sC1[bufr / kBwCols][bufr % kBwCols][c] = 0.f;
    if (kKeyRows || !v_t) sC2[bufr / kBwCols][(bufr % kBwCols) * PITCH + c] = 0.f;
  }
  if (!kKeyRows && v_t)  // transposed V block: the padded depth rows D..DP-1
    for (int i = tid; i < 2 * (DP - D) * PT; i += kBwThreads) sC2[i / ((DP - D) * PT)][D * PT + i % ((DP - D) * PT)] = 0.f;
  issue(0, 0);

  // ---- this warp's 16 rows: A fragments of X1 (q or k) and X2 (dctx or v), kept for the whole kernel -------------
  const int rA = row0 + g, rB = row0 + g + 8;
  const bool vA = rA < Nrows, vB = rB < Nrows;
  uint32_t x1f[KS][4], x2f[KS][4];
  float dA = 0.f, dB = 0.f;  // delta of rows rA / rB (rows = queries)
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = (e & 1) ? rB : rA;
      const int d = ks * 8 + t + ((e & 2) ? 4 : 0);
      const bool ok = ((e & 1) ? vB : vA) && d < D;
      const size_t off = (size_t)row * HD + h * D + d;
      const float a = ok ? __ldg(x1base + off) : 0.f;
      const float c = ok ? ((kKeyRows && v_t) ? __ldg(vtb + (size_t)d * p.ldv + row) : __ldg(x2base + off)) : 0.f;
      x1f[ks][e] = f2tf32(a);
      x2f[ks][e] = f2tf32(c);
      if (!kKeyRows && ok) {
        const float o = __ldg(p.ctx + (size_t)b * p.Nq * HD + off);
        if (e & 1) dB = fmaf(c, o, dB); else dA = fmaf(c, o, dA);
      }
    }
  }
  float lseA = 0.f, lseB = 0.f, maskA = 0.f, maskB = 0.f;
  if (!kKeyRows) {
    dA += __shfl_xor_sync(kFullMask, dA, 1); dA += __shfl_xor_sync(kFullMask, dA, 2);
    dB += __shfl_xor_sync(kFullMask, dB, 1); dB += __shfl_xor_sync(kFullMask, dB, 2);
    if (t == 0) {
      if (vA) p.delta[stat_base + rA] = dA;
      if (vB) p.delta[stat_base + rB] = dB;
    }
    lseA = vA ? __ldg(p.lse + stat_base + rA) * kLog2e : INFINITY;
    lseB = vB ? __ldg(p.lse + stat_base + rB) * kLog2e : INFINITY;
  } else {
    // rows = keys: the key-padding mask is a property of the row
    bool kA = vA, kB = vB;
    if (p.mask) {
      if (kA) kA = p.mask[(size_t)b * p.Nk + rA] == 0;
      if (kB) kB = p.mask[(size_t)b * p.Nk + rB] == 0;
    }
    maskA = kA ? 0.f : -INFINITY;
    maskB = kB ? 0.f : -INFINITY;
  }

  float o1[KS][4], o2[KS][4];  // o1: dq or dk; o2: dv (rows = keys only)
#pragma unroll
  for (int i = 0; i < KS; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { o1[i][e] = 0.f; o2[i][e] = 0.f; }

  const float sl2 = p.scale * kLog2e;
  const uint32_t dseed = kDrop ? effective_seed(p.drop_seed, p.seed_epoch) : 0u;
  for (int blk = 0; blk < nblocks; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < nblocks) {
      issue(blk + 1, buf ^ 1);
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
    const int c0 = blk * kBwCols;
    // ---- S = X1 C1^T, dP = X2 C2^T -----------------------------------------------------------------------------
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; dp[j][e] = 0.f; }
      const float *c1r = &sC1[buf][j * 8 + g][t];
      // V block: [key][PITCH] (element (key, dim) at key * PITCH + dim) or transposed [dim][PT]
      const bool c2_t = !kKeyRows && v_t;
      const float *c2r = c2_t ? &sC2[buf][t * PT + j * 8 + g] : &sC2[buf][(j * 8 + g) * PITCH + t];
      const int c2s = c2_t ? PT : 1;  // stride of one depth step
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        mma_m16n8k8_tf32(s[j], x1f[ks], f2tf32(c1r[ks * 8]), f2tf32(c1r[ks * 8 + 4]));
        mma_m16n8k8_tf32(dp[j], x2f[ks], f2tf32(c2r[(ks * 8) * c2s]), f2tf32(c2r[(ks * 8 + 4) * c2s]));
      }
    }
    // ---- P, dS (element e of tile j: row (e & 2 ? rB : rA), column c0 + 8 j + 2 t + (e & 1)) ------------------------
    uint32_t dsf[8][4], pf[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = j * 8 + 2 * t + (e & 1);
        const bool second = (e & 2) != 0;
        float l2, dl, madd;
        if (kKeyRows) {
          l2 = sStat[buf][0][cl];
          dl = sStat[buf][1][cl];
          madd = second ? maskB : maskA;
        } else {
          l2 = second ? lseB : lseA;
          dl = second ? dB : dA;
          madd = sStat[buf][0][cl];
        }
        float pe = ex2_approx(fmaf(s[j][e], sl2, madd) - l2);
        float dpe = dp[j][e];
        if (kDrop) {
          const int qi = kKeyRows ? c0 + cl : (second ? rB : rA);
          const int ki = kKeyRows ? (second ? rB : rA) : c0 + cl;
          const bool keep = dropout_keep(dseed, (uint32_t)(stat_base + qi), (uint32_t)ki, p.drop_thresh);
          dpe = keep ? dpe * p.drop_scale : 0.f;
          pf[j][e] = f2tf32(keep ? pe * p.drop_scale : 0.f);
        } else {
          pf[j][e] = f2tf32(pe);
        }
        dsf[j][e] = f2tf32(pe * (dpe - dl));
      }
    }
    // ---- o1 += dS C1 (and o2 += P C2): k index = column, permuted so that accumulator columns (2t, 2t+1) are the
    // fragment's (t, t+4): A = {e0, e2, e1, e3}, B rows 8 j + 2 t and 8 j + 2 t + 1 --------------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a1[4] = {dsf[j][0], dsf[j][2], dsf[j][1], dsf[j][3]};
      const uint32_t a2[4] = {pf[j][0], pf[j][2], pf[j][1], pf[j][3]};
      const float *c1r = &sC1[buf][j * 8 + 2 * t][g], *c2r = &sC2[buf][(j * 8 + 2 * t) * PITCH + g];
#pragma unroll
      for (int nt = 0; nt < KS; ++nt) {
        mma_m16n8k8_tf32(o1[nt], a1, f2tf32(c1r[nt * 8]), f2tf32(c1r[PITCH + nt * 8]));
        if (kKeyRows) mma_m16n8k8_tf32(o2[nt], a2, f2tf32(c2r[nt * 8]), f2tf32(c2r[PITCH + nt * 8]));
      }
    }
    __syncthreads();  // the next iteration's prefetch overwrites the other buffer, which this one just read... (see below)
  }

  // ---- store: element e of n-tile nt: row (e & 2 ? rB : rA), depth 8 nt + 2 t + (e & 1) ---------------------------
  float *out1 = (kKeyRows ? p.dk + (size_t)b * p.Nk * HD : p.dq + (size_t)b * p.Nq * HD) + h * D;
  float *out2 = kKeyRows ? p.dv + (size_t)b * p.Nk * HD + h * D : nullptr;
#pragma unroll
  for (int nt = 0; nt < KS; ++nt) {
    const int d = nt * 8 + 2 * t;
    if (d < D) {  // D is even: d + 1 < D as well
      if (vA) {
        *reinterpret_cast<float2 *>(out1 + (size_t)rA * HD + d) = make_float2(o1[nt][0] * p.scale, o1[nt][1] * p.scale);
        if (kKeyRows) *reinterpret_cast<float2 *>(out2 + (size_t)rA * HD + d) = make_float2(o2[nt][0], o2[nt][1]);
      }
      if (vB) {
        *reinterpret_cast<float2 *>(out1 + (size_t)rB * HD + d) = make_float2(o1[nt][2] * p.scale, o1[nt][3] * p.scale);
        if (kKeyRows) *reinterpret_cast<float2 *>(out2 + (size_t)rB * HD + d) = make_float2(o2[nt][2], o2[nt][3]);
      }
    }
  }
}

template <int D>
int launch_attention_backward(const AttnBwdParams &p, int B, cudaStream_t st) {
  dim3 gq((unsigned)((p.Nq + kBwRows - 1) / kBwRows), (unsigned)p.H, (unsigned)B);
  dim3 gk((unsigned)((p.Nk + kBwRows - 1) / kBwRows), (unsigned)p.H, (unsigned)B);
  if (p.drop_thresh) {
    EDA_CUDA_TRY(launch_pdl(attention_backward_kernel<D, false, true>, gq, dim3(kBwThreads), 0, st, p), "attention backward launch");
    EDA_CUDA_TRY(launch_pdl(attention_backward_kernel<D, true, true>, gk, dim3(kBwThreads), 0, st, p), "attention backward launch");
  } else {
    EDA_CUDA_TRY(launch_pdl(attention_backward_kernel<D, false, false>, gq, dim3(kBwThreads), 0, st, p), "attention backward launch");
    EDA_CUDA_TRY(launch_pdl(attention_backward_kernel<D, true, false>, gk, dim3(kBwThreads), 0, st, p), "attention backward launch");
  }
  return check_launch("attention_backward_kernel", 2);
}

}  // namespace
}  // namespace eda

extern "C" int eda_attention_backward(const float *q, const float *k, const float *v, long long v_batch_stride,
                                      const float *vt, int ldv, const float *dctx, const float *ctx, const float *lse,
                                      const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                      float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *delta, float *dq,
                                      float *dk, float *dv, void *stream) {
  using namespace eda;
  if (B < 0 || Nq < 0 || Nk < 1 || H < 1 || D < 1) return EDA_ERR_INVALID_ARGUMENT;
  if (H > 65535 || B > 65535) return EDA_ERR_UNSUPPORTED;
  if (B == 0 || Nq == 0) return EDA_OK;
  if (!q || !k || (!v && !vt) || !dctx || !ctx || !lse || !delta || !dq || !dk || !dv) return EDA_ERR_INVALID_ARGUMENT;
  if (vt) {
    if (ldv < Nk || (ldv & 3)) return EDA_ERR_INVALID_ARGUMENT;
  } else if (v_batch_stride < (long long)Nk * H * D || (v_batch_stride & 3)) {
    return EDA_ERR_INVALID_ARGUMENT;
  }
  const void *ptrs[] = {q, k, vt ? vt : v, dctx, ctx, dq, dk, dv};
  for (const void *ptr : ptrs)
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return EDA_ERR_INVALID_ARGUMENT;
  if (dropout_p < 0.f || dropout_p >= 1.f) return EDA_ERR_INVALID_ARGUMENT;
  AttnBwdParams p = {};
  p.q = q; p.k = k; p.v = v; p.dctx = dctx; p.ctx = ctx; p.lse = lse; p.mask = key_padding_mask; p.delta = delta;
  p.dq = dq; p.dk = dk; p.dv = dv; p.v_batch_stride = v_batch_stride; p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.vt = vt; p.ldv = ldv;
  if (vt) p.v = vt;  // never dereferenced in that mode; keeps the pointer arithmetic defined
  p.drop_thresh = dropout_thresh(dropout_p); p.drop_seed = dropout_seed; p.drop_scale = 1.0f / (1.0f - dropout_p);
  p.seed_epoch = reinterpret_cast<const uint32_t *>(dropout_epoch);
  cudaStream_t st = as_stream(stream);
  switch (D) {
    case 32: return launch_attention_backward<32>(p, B, st);
    case 36: return launch_attention_backward<36>(p, B, st);
    default: return EDA_ERR_UNSUPPORTED;
  }
}
