// Backward of the fused set-abstraction stage (neighbour gather + 3 x [1x1 conv -> BatchNorm -> ReLU] + max-pool,
// csrc/sa_mlp.cu) for sm_100a.
//
// In the reference this is autograd through QueryAndGroup / SharedMLP / max_pool2d
// (pointnet2/pointnet2_utils.py:209-257,317-376, pointnet2/pytorch_utils.py:11-36, pointnet2_modules.py:251-267):
// cuDNN convolution + BatchNorm backward over (B,C,npoint,nsample) tensors and the atomic scatter of
// group_points_grad (group_points_gpu.cu:48-80).  Here the grouped rows are row-major (row = grouped point, column =
// channel) so that every layer is a plain GEMM on the forward tcgen05 kernel (eda_linear_forward) and the split-row
// weight-gradient kernel (eda_wgrad); this file holds the memory-bound stages in between, each one pass over HBM:
//
//   eda_sa_gather_rows            x0[r] = [features[idx[r]] | (xyz[idx[r]] - centre) (/ radius) | 0]   (layer-1 input)
//   eda_bn_relu_apply             a = relu(z * scale[c] + shift[c])
//   eda_sa_pool_backward          max-pool + last ReLU + BatchNorm-3 reductions: per (centre, channel) the first arg-max
//                                 row, sum(dy), sum(dy * zhat)
//   eda_sa_pool_backward_apply    dz3 = scale3 (dy3 - mean(dy3) - zhat3 mean(dy3 zhat3)), in place over z3
//   eda_bn_relu_backward_stats    sum(dy), sum(dy * zhat) with dy = da * [z scale + shift > 0]
//   eda_bn_relu_backward_apply    dz = scale (dy - mean(dy) - zhat mean(dy zhat)), in place over da
//   eda_sa_scatter_rows           d features[idx[r]] += dx0[r][:C]      (red.global.add.f32, like the reference)
//
// sum(dy) and sum(dy * zhat) are also the gradients of the BatchNorm bias and weight.
#include "common.cuh"

namespace eda {
namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__global__ void __launch_bounds__(256)
sa_gather_rows_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, const float *__restrict__ feat,
                      int feat_stride, const int *__restrict__ idx, long long total_rows, int N, int M, int S, int C,
                      int K0pad, float radius, int normalize, float *__restrict__ x0) {
  const int nch = K0pad >> 2;
  const long long total = total_rows * nch;
  const bool vec_ok = (feat_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
  for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
       id += (long long)gridDim.x * blockDim.x) {
    const long long row = id / nch;
    const int c = (int)(id - row * nch) << 2;
    const long long bj = row / S;  // b * M + j
    const long long b = bj / M;
    const int pi = __ldg(idx + row);
    float4 o;
    if (c + 4 <= C && vec_ok) {
      o = __ldg(reinterpret_cast<const float4 *>(feat + (b * N + pi) * (long long)feat_stride + c));
    } else {
      float f[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = c + e;
        float v = 0.f;
        if (cc < C) {
          v = __ldg(feat + (b * N + pi) * (long long)feat_stride + cc);
        } else if (cc < C + 3) {
          const int a = cc - C;
          v = __fsub_rn(__ldg(xyz + (b * N + pi) * 3 + a), __ldg(new_xyz + bj * 3 + a));  // pointnet2_utils.py:350
          if (normalize) v = __fdiv_rn(v, radius);                                        // :351-352
        }
        f[e] = v;
      }
      o = make_float4(f[0], f[1], f[2], f[3]);
    }
    reinterpret_cast<float4 *>(x0)[id] = o;
  }
}

__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const float4 *__restrict__ z, const float *__restrict__ scale, const float *__restrict__ shift,
                     long long n4, int C, float4 *__restrict__ out) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) << 2;
    const float4 v = __ldg(z + i);
    const float4 s = __ldg(reinterpret_cast<const float4 *>(scale + c));
    const float4 t = __ldg(reinterpret_cast<const float4 *>(shift + c));
    out[i] = make_float4(fmaxf(fmaf(v.x, s.x, t.x), 0.f), fmaxf(fmaf(v.y, s.y, t.y), 0.f),
                         fmaxf(fmaf(v.z, s.z, t.z), 0.f), fmaxf(fmaf(v.w, s.w, t.w), 0.f));
  }
}

// thread = channel, block loops over centres: arg-max over the S rows of a centre, ReLU gate, BN reductions
__global__ void __launch_bounds__(256)
sa_pool_backward_kernel(const float *__restrict__ z3, const float *__restrict__ scale, const float *__restrict__ shift,
                        const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gout,
                        long long centres, int S, int C, int *__restrict__ amax, float *__restrict__ stats) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float sc = __ldg(scale + c), sh = __ldg(shift + c), mu = __ldg(mean + c), is = __ldg(invstd + c);
  float s1 = 0.f, s2 = 0.f;
  for (long long j = blockIdx.x; j < centres; j += gridDim.x) {
    const float *zr = z3 + j * S * (long long)C + c;
    float best = -INFINITY, zbest = 0.f;
    int bi = 0;
    for (int s = 0; s < S; ++s) {
      const float zv = __ldg(zr + (long long)s * C);
      const float y = fmaf(zv, sc, sh);
      if (y > best) { best = y; bi = s; zbest = zv; }
    }
    const bool on = best > 0.f;
    amax[j * C + c] = on ? bi : -1;
    if (on) {
      const float dy = __ldg(gout + j * C + c);
      s1 += dy;
      s2 = fmaf(dy, (zbest - mu) * is, s2);
    }
  }
  atomicAdd(stats + c, s1);
  atomicAdd(stats + C + c, s2);
}

__global__ void __launch_bounds__(256)
sa_pool_backward_apply_kernel(float4 *__restrict__ z3, const int *__restrict__ amax, const float *__restrict__ gout,
                              const float *__restrict__ scale, const float *__restrict__ mean,
                              const float *__restrict__ invstd, const float *__restrict__ stats, float inv_count,
                              int batch_stats, long long n4, int S, int C) {
  const int c4n = C >> 2;
  const bool small = n4 < 0x7fffffffLL;  // 32-bit index arithmetic (64-bit divisions are ~100 instructions each)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    long long row, j;
    int c, s;
    if (small) {
      const unsigned iu = (unsigned)i, ru = iu / (unsigned)c4n, ju = ru / (unsigned)S;
      row = ru; j = ju; c = (int)(iu - ru * (unsigned)c4n) << 2; s = (int)(ru - ju * (unsigned)S);
    } else {
      row = i / c4n; c = (int)(i - row * c4n) << 2; j = row / S; s = (int)(row - j * S);
    }
    const float4 z = z3[i];
    const int4 am = __ldg(reinterpret_cast<const int4 *>(amax + j * C + c));
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gout + j * C + c));
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
    float4 o;
    const float dy0 = am.x == s ? g.x : 0.f, dy1 = am.y == s ? g.y : 0.f, dy2 = am.z == s ? g.z : 0.f,
                dy3 = am.w == s ? g.w : 0.f;
    if (batch_stats) {
      const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
      const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
      const float4 m1 = __ldg(reinterpret_cast<const float4 *>(stats + c));
      const float4 m2 = __ldg(reinterpret_cast<const float4 *>(stats + C + c));
      o.x = sc.x * (dy0 - m1.x * inv_count - (z.x - mu.x) * is.x * (m2.x * inv_count));
      o.y = sc.y * (dy1 - m1.y * inv_count - (z.y - mu.y) * is.y * (m2.y * inv_count));
      o.z = sc.z * (dy2 - m1.z * inv_count - (z.z - mu.z) * is.z * (m2.z * inv_count));
      o.w = sc.w * (dy3 - m1.w * inv_count - (z.w - mu.w) * is.w * (m2.w * inv_count));
    } else {
      o = make_float4(sc.x * dy0, sc.y * dy1, sc.z * dy2, sc.w * dy3);
    }
    z3[i] = o;
  }
}

// column sums over (rows, C): thread = (row lane, float4 column)
__global__ void __launch_bounds__(256)
bn_relu_backward_stats_kernel(const float4 *__restrict__ da, const float4 *__restrict__ z, const float *__restrict__ scale,
                              const float *__restrict__ shift, const float *__restrict__ mean,
                              const float *__restrict__ invstd, long long rows, int C, float *__restrict__ stats) {
  __shared__ float s_acc[2 * 512];
  const int c4n = C >> 2;
  const int col = threadIdx.x % c4n, lane_r = threadIdx.x / c4n, rstep = blockDim.x / c4n;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int c = col << 2;
  const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
  const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (lane_r < rstep) {
    for (long long r = (long long)blockIdx.x * rstep + lane_r; r < rows; r += (long long)gridDim.x * rstep) {
      const float4 d = __ldg(da + r * c4n + col), zv = __ldg(z + r * c4n + col);
      const float d0 = fmaf(zv.x, sc.x, sh.x) > 0.f ? d.x : 0.f, d1 = fmaf(zv.y, sc.y, sh.y) > 0.f ? d.y : 0.f,
                  d2 = fmaf(zv.z, sc.z, sh.z) > 0.f ? d.z : 0.f, d3 = fmaf(zv.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
      a1.x += d0; a1.y += d1; a1.z += d2; a1.w += d3;
      a2.x = fmaf(d0, (zv.x - mu.x) * is.x, a2.x); a2.y = fmaf(d1, (zv.y - mu.y) * is.y, a2.y);
      a2.z = fmaf(d2, (zv.z - mu.z) * is.z, a2.z); a2.w = fmaf(d3, (zv.w - mu.w) * is.w, a2.w);
    }
    atomicAdd(&s_acc[c + 0], a1.x); atomicAdd(&s_acc[c + 1], a1.y); atomicAdd(&s_acc[c + 2], a1.z); atomicAdd(&s_acc[c + 3], a1.w);
    atomicAdd(&s_acc[C + c + 0], a2.x); atomicAdd(&s_acc[C + c + 1], a2.y);
    atomicAdd(&s_acc[C + c + 2], a2.z); atomicAdd(&s_acc[C + c + 3], a2.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(stats + i, s_acc[i]);
}

__global__ void __launch_bounds__(256)
bn_relu_backward_apply_kernel(float4 *__restrict__ da, const float4 *__restrict__ z, const float *__restrict__ scale,
                              const float *__restrict__ shift, const float *__restrict__ mean,
                              const float *__restrict__ invstd, const float *__restrict__ stats, float inv_count,
                              int batch_stats, long long n4, int C) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) << 2;
    const float4 d = da[i], zv = __ldg(z + i);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
    const float d0 = fmaf(zv.x, sc.x, sh.x) > 0.f ? d.x : 0.f, d1 = fmaf(zv.y, sc.y, sh.y) > 0.f ? d.y : 0.f,
                d2 = fmaf(zv.z, sc.z, sh.z) > 0.f ? d.z : 0.f, d3 = fmaf(zv.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
    float4 o;
    if (batch_stats) {
      const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
      const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
      const float4 m1 = __ldg(reinterpret_cast<const float4 *>(stats + c));
      const float4 m2 = __ldg(reinterpret_cast<const float4 *>(stats + C + c));
      o.x = sc.x * (d0 - m1.x * inv_count - (zv.x - mu.x) * is.x * (m2.x * inv_count));
      o.y = sc.y * (d1 - m1.y * inv_count - (zv.y - mu.y) * is.y * (m2.y * inv_count));
      o.z = sc.z * (d2 - m1.z * inv_count - (zv.z - mu.z) * is.z * (m2.z * inv_count));
      o.w = sc.w * (d3 - m1.w * inv_count - (zv.w - mu.w) * is.w * (m2.w * inv_count));
    } else {
      o = make_float4(sc.x * d0, sc.y * d1, sc.z * d2, sc.w * d3);
    }
    da[i] = o;
  }
}

// per-channel (sum, sum of squares) of z (rows, C) -> fp64 accumulators (train-mode BatchNorm statistics of the
// row-major training forward)
__global__ void __launch_bounds__(256)
col_stats_kernel(const float4 *__restrict__ z, long long rows, int C, double *__restrict__ stats) {
  // Every thread owns a fixed set of rows and 4 columns; the row lanes of a column are then summed in a fixed order and
  // each block contributes ONE fp64 atomic per entry, so the result does not depend on scheduling beyond fp64 rounding
  // (train-mode outputs stay reproducible from run to run; see csrc/sa_mlp.cu for why that matters with tf32).
  __shared__ float s_part[2048];  // [row lane][2 C]: rstep * 2 C = 2048 floats for every C
  const int c4n = C >> 2;
  const int col = threadIdx.x % c4n, lane_r = threadIdx.x / c4n, rstep = blockDim.x / c4n;
  const int c = col << 2;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (lane_r < rstep) {
    for (long long r = (long long)blockIdx.x * rstep + lane_r; r < rows; r += (long long)gridDim.x * rstep) {
      const float4 v = __ldg(z + r * c4n + col);
      a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
      a2.x = fmaf(v.x, v.x, a2.x); a2.y = fmaf(v.y, v.y, a2.y); a2.z = fmaf(v.z, v.z, a2.z); a2.w = fmaf(v.w, v.w, a2.w);
    }
    float *dst = s_part + (size_t)lane_r * 2 * C;
    dst[c + 0] = a1.x; dst[c + 1] = a1.y; dst[c + 2] = a1.z; dst[c + 3] = a1.w;
    dst[C + c + 0] = a2.x; dst[C + c + 1] = a2.y; dst[C + c + 2] = a2.z; dst[C + c + 3] = a2.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double acc = 0.0;
    for (int l = 0; l < rstep; ++l) acc += (double)s_part[(size_t)l * 2 * C + i];
    atomicAdd(stats + i, acc);
  }
}

// out[j][c] = max(0, max_s (z3[j*S + s][c] * scale[c] + shift[c])): BatchNorm-3 + ReLU + max-pool of the row-major
// training forward; thread = channel, block loops over centres
__global__ void __launch_bounds__(256)
sa_pool_forward_kernel(const float *__restrict__ z3, const float *__restrict__ scale, const float *__restrict__ shift,
                       long long centres, int S, int C, float *__restrict__ out, int *__restrict__ amax) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float sc = __ldg(scale + c), sh = __ldg(shift + c);
  for (long long j = blockIdx.x; j < centres; j += gridDim.x) {
    const float *zr = z3 + j * S * (long long)C + c;
    float best = 0.f;
    int bi = -1;  // first row that attains the maximum, -1 when nothing is positive (same rule as sa_pool_backward_kernel)
    for (int s = 0; s < S; ++s) {
      const float y = fmaf(__ldg(zr + (long long)s * C), sc, sh);
      if (y > best) { best = y; bi = s; }
    }
    out[j * C + c] = best;
    if (amax) amax[j * C + c] = bi;
  }
}

// Same result with thread = (centre slot, 4 channels): 16-byte loads, 256 / (C / 4) centres per block in flight and the S
// loads of a thread independent of each other — the thread = channel form above keeps 4-byte loads and, for C = 128, half
// of every block idle (193 us for SA1's 537 MB, 2.8 TB/s).  C in {64, 128, 256}.
__global__ void __launch_bounds__(256)
sa_pool_forward_vec_kernel(const float *__restrict__ z3, const float *__restrict__ scale, const float *__restrict__ shift,
                           long long centres, int S, int C, float *__restrict__ out, int *__restrict__ amax) {
  const int c4n = C >> 2, cpb = 256 / c4n;
  const int slot = threadIdx.x / c4n, c = (threadIdx.x - slot * c4n) << 2;
  const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
  for (long long j = (long long)blockIdx.x * cpb + slot; j < centres; j += (long long)gridDim.x * cpb) {
    const float4 *zr = reinterpret_cast<const float4 *>(z3 + j * S * (long long)C + c);
    float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 bi = make_int4(-1, -1, -1, -1);  // first row that attains the maximum, -1 when nothing is positive
    for (int s0 = 0; s0 < S; s0 += 8) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = s0 + i < S ? __ldg(zr + (long long)(s0 + i) * c4n) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y0 = fmaf(v[i].x, sc.x, sh.x), y1 = fmaf(v[i].y, sc.y, sh.y);
        const float y2 = fmaf(v[i].z, sc.z, sh.z), y3 = fmaf(v[i].w, sc.w, sh.w);
        if (s0 + i < S) {
          if (y0 > best.x) { best.x = y0; bi.x = s0 + i; }
          if (y1 > best.y) { best.y = y1; bi.y = s0 + i; }
          if (y2 > best.z) { best.z = y2; bi.z = s0 + i; }
          if (y3 > best.w) { best.w = y3; bi.w = s0 + i; }
        }
      }
    }
    *reinterpret_cast<float4 *>(out + j * C + c) = best;
    if (amax) *reinterpret_cast<int4 *>(amax + j * C + c) = bi;
  }
}

// BatchNorm-3 reductions from a saved arg-max: stats[0:C] += sum dy, stats[C:2C] += sum dy * zhat, dy = grad_out at amax
__global__ void __launch_bounds__(256)
sa_pool_backward_stats_kernel(const float *__restrict__ z3, const int *__restrict__ amax, const float *__restrict__ mean,
                              const float *__restrict__ invstd, const float *__restrict__ gout, long long centres, int S,
                              int C, float *__restrict__ stats) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c);
  float s1 = 0.f, s2 = 0.f;
  for (long long j = blockIdx.x; j < centres; j += gridDim.x) {
    const int bi = __ldg(amax + j * C + c);
    if (bi >= 0) {
      const float dy = __ldg(gout + j * C + c);
      const float zv = __ldg(z3 + (j * S + bi) * (long long)C + c);
      s1 += dy;
      s2 = fmaf(dy, (zv - mu) * is, s2);
    }
  }
  atomicAdd(stats + c, s1);
  atomicAdd(stats + C + c, s2);
}

__global__ void __launch_bounds__(256)
sa_scatter_rows_kernel(const float *__restrict__ dx0, const int *__restrict__ idx, long long total_rows, int N, int MS,
                       int C, int K0pad, float *__restrict__ dfeat) {
  const int nch = (C + 3) >> 2;
  const long long total = total_rows * nch;
  for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total;
       id += (long long)gridDim.x * blockDim.x) {
    const long long row = id / nch;
    const int c = (int)(id - row * nch) << 2;
    const long long b = row / MS;
    const int pi = __ldg(idx + row);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(dx0 + row * K0pad + c));
    float *dst = dfeat + (b * N + pi) * (long long)C + c;
    atomicAdd(dst, v.x);
    if (c + 1 < C) atomicAdd(dst + 1, v.y);
    if (c + 2 < C) atomicAdd(dst + 2, v.z);
    if (c + 3 < C) atomicAdd(dst + 3, v.w);
  }
}

inline unsigned grid_for(long long work_items, int per_block = 256, int max_blocks = 148 * 16) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}
inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace eda

extern "C" {

int eda_sa_gather_rows(const float *xyz, const float *new_xyz, const float *feat, int feat_stride, const int *idx, int B,
                       int N, int M, int S, int C, int K0pad, float radius, int normalize_xyz, float *x0, void *stream) {
  using namespace eda;
  if (B < 0 || N <= 0 || M < 0 || S <= 0 || C < 0 || K0pad < C + 3 || (K0pad & 3)) return EDA_ERR_INVALID_ARGUMENT;
  const long long rows = (long long)B * M * S;
  if (rows == 0) return EDA_OK;
  if (!xyz || !new_xyz || !idx || !x0 || (C > 0 && (!feat || feat_stride < C)) || !al16(x0)) return EDA_ERR_INVALID_ARGUMENT;
  sa_gather_rows_kernel<<<grid_for(rows * (K0pad >> 2)), 256, 0, as_stream(stream)>>>(
      xyz, new_xyz, feat, feat_stride, idx, rows, N, M, S, C, K0pad, radius, normalize_xyz, x0);
  return check_launch("sa_gather_rows_kernel");
}

int eda_bn_relu_apply(const float *z, const float *scale, const float *shift, long long rows, int C, float *out,
                      void *stream) {
  using namespace eda;
  if (rows < 0 || C < 4 || (C & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0) return EDA_OK;
  if (!z || !scale || !shift || !out || !al16(z) || !al16(out) || !al16(scale) || !al16(shift)) return EDA_ERR_INVALID_ARGUMENT;
  const long long n4 = rows * (C >> 2);
  bn_relu_apply_kernel<<<grid_for(n4), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(z), scale, shift, n4, C,
                                                                   reinterpret_cast<float4 *>(out));
  return check_launch("bn_relu_apply_kernel");
}

int eda_sa_pool_backward(const float *z3, const float *scale, const float *shift, const float *mean, const float *invstd,
                         const float *grad_out, long long centres, int S, int C, int *amax, float *stats, void *stream) {
  using namespace eda;
  if (centres < 0 || S <= 0 || C < 4 || (C & 3) || C > 256) return EDA_ERR_INVALID_ARGUMENT;
  if (centres == 0) return EDA_OK;
  if (!z3 || !scale || !shift || !mean || !invstd || !grad_out || !amax || !stats) return EDA_ERR_INVALID_ARGUMENT;
  sa_pool_backward_kernel<<<grid_for(centres, 1, 148 * 8), 256, 0, as_stream(stream)>>>(z3, scale, shift, mean, invstd,
                                                                                    grad_out, centres, S, C, amax, stats);
  return check_launch("sa_pool_backward_kernel");
}

int eda_sa_pool_backward_apply(float *z3, const int *amax, const float *grad_out, const float *scale, const float *mean,
                               const float *invstd, const float *stats, double count, int batch_stats, long long centres,
                               int S, int C, void *stream) {
  using namespace eda;
  if (centres < 0 || S <= 0 || C < 4 || (C & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (centres == 0) return EDA_OK;
  if (!z3 || !amax || !grad_out || !scale || !al16(z3) || !al16(amax) || !al16(grad_out) || !al16(scale))
    return EDA_ERR_INVALID_ARGUMENT;
  if (batch_stats && (!mean || !invstd || !stats || count <= 0 || !al16(mean) || !al16(invstd) || !al16(stats)))
    return EDA_ERR_INVALID_ARGUMENT;
  const long long n4 = centres * S * (C >> 2);
  sa_pool_backward_apply_kernel<<<grid_for(n4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4 *>(z3), amax, grad_out, scale, mean, invstd, stats, batch_stats ? (float)(1.0 / count) : 0.f,
      batch_stats, n4, S, C);
  return check_launch("sa_pool_backward_apply_kernel");
}

int eda_bn_relu_backward_stats(const float *da, const float *z, const float *scale, const float *shift, const float *mean,
                               const float *invstd, long long rows, int C, float *stats, void *stream) {
  using namespace eda;
  if (rows < 0 || C < 4 || (C & 3) || C > 512) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0) return EDA_OK;
  if (!da || !z || !scale || !shift || !mean || !invstd || !stats || !al16(da) || !al16(z) || !al16(scale) || !al16(shift) ||
      !al16(mean) || !al16(invstd))
    return EDA_ERR_INVALID_ARGUMENT;
  const int rstep = 256 / (C >> 2);
  bn_relu_backward_stats_kernel<<<grid_for(rows, rstep * 8, 148 * 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(da), reinterpret_cast<const float4 *>(z), scale, shift, mean, invstd, rows, C, stats);
  return check_launch("bn_relu_backward_stats_kernel");
}

int eda_bn_relu_backward_apply(float *da, const float *z, const float *scale, const float *shift, const float *mean,
                               const float *invstd, const float *stats, double count, int batch_stats, long long rows,
                               int C, void *stream) {
  using namespace eda;
  if (rows < 0 || C < 4 || (C & 3)) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0) return EDA_OK;
  if (!da || !z || !scale || !shift || !al16(da) || !al16(z) || !al16(scale) || !al16(shift)) return EDA_ERR_INVALID_ARGUMENT;
  if (batch_stats && (!mean || !invstd || !stats || count <= 0 || !al16(mean) || !al16(invstd) || !al16(stats)))
    return EDA_ERR_INVALID_ARGUMENT;
  const long long n4 = rows * (C >> 2);
  bn_relu_backward_apply_kernel<<<grid_for(n4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4 *>(da), reinterpret_cast<const float4 *>(z), scale, shift, mean, invstd, stats,
      batch_stats ? (float)(1.0 / count) : 0.f, batch_stats, n4, C);
  return check_launch("bn_relu_backward_apply_kernel");
}

int eda_col_stats(const float *z, long long rows, int C, double *stats, void *stream) {
  using namespace eda;
  if (rows < 0 || C < 4 || (C & 3) || C > 512) return EDA_ERR_INVALID_ARGUMENT;
  if (rows == 0) return EDA_OK;
  if (!z || !stats || !al16(z)) return EDA_ERR_INVALID_ARGUMENT;
  const int rstep = 256 / (C >> 2);
  col_stats_kernel<<<grid_for(rows, rstep * 8, 148 * 8), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(z),
                                                                                   rows, C, stats);
  return check_launch("col_stats_kernel");
}

int eda_sa_pool_backward_stats(const float *z3, const int *amax, const float *mean, const float *invstd,
                               const float *grad_out, long long centres, int S, int C, float *stats, void *stream) {
  using namespace eda;
  if (centres < 0 || S <= 0 || C < 1 || C > 256) return EDA_ERR_INVALID_ARGUMENT;
  if (centres == 0) return EDA_OK;
  if (!z3 || !amax || !mean || !invstd || !grad_out || !stats) return EDA_ERR_INVALID_ARGUMENT;
  sa_pool_backward_stats_kernel<<<grid_for(centres, 4, 148 * 8), 256, 0, as_stream(stream)>>>(z3, amax, mean, invstd, grad_out,
                                                                                          centres, S, C, stats);
  return check_launch("sa_pool_backward_stats_kernel");
}

int eda_sa_pool_forward(const float *z3, const float *scale, const float *shift, long long centres, int S, int C,
                        float *out, int *amax, void *stream) {
  using namespace eda;
  if (centres < 0 || S <= 0 || C < 1 || C > 256) return EDA_ERR_INVALID_ARGUMENT;
  if (centres == 0) return EDA_OK;
  if (!z3 || !scale || !shift || !out) return EDA_ERR_INVALID_ARGUMENT;
  if ((C == 64 || C == 128 || C == 256) && al16(z3) && al16(scale) && al16(shift) && al16(out) && (!amax || al16(amax))) {
    const int cpb = 256 / (C >> 2);
    sa_pool_forward_vec_kernel<<<grid_for(centres, cpb, 148 * 8), 256, 0, as_stream(stream)>>>(z3, scale, shift, centres, S, C,
                                                                                           out, amax);
    return check_launch("sa_pool_forward_vec_kernel");
  }
  sa_pool_forward_kernel<<<grid_for(centres, 1, 148 * 8), 256, 0, as_stream(stream)>>>(z3, scale, shift, centres, S, C, out,
                                                                                   amax);
  return check_launch("sa_pool_forward_kernel");
}

int eda_sa_scatter_rows(const float *dx0, const int *idx, int B, int N, int M, int S, int C, int K0pad, float *dfeat,
                        void *stream) {
  using namespace eda;
  if (B < 0 || N <= 0 || M < 0 || S <= 0 || C < 1 || K0pad < C || (K0pad & 3)) return EDA_ERR_INVALID_ARGUMENT;
  const long long rows = (long long)B * M * S;
  if (rows == 0) return EDA_OK;
  if (!dx0 || !idx || !dfeat || !al16(dx0)) return EDA_ERR_INVALID_ARGUMENT;
  sa_scatter_rows_kernel<<<grid_for(rows * ((C + 3) >> 2)), 256, 0, as_stream(stream)>>>(dx0, idx, rows, N, M * S, C, K0pad,
                                                                                     dfeat);
  return check_launch("sa_scatter_rows_kernel");
}

}  // extern "C"
