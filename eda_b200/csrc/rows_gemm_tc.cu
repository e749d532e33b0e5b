// Row-streaming GEMM of the set-abstraction stage on the tcgen05 tensor cores (sm_100a), persistent:
//
//     Y[R x N] = f(X)[R x K] W'[N x K]^T,   f(x) = x   or   relu(x * in_scale[k] + in_shift[k])
//     (+ optionally the column sums / sums of squares of Y: the BatchNorm batch statistics of the layer)
//
// Same contract as the warp-level kernel of rows_gemm.cu (eda_rows_gemm / eda_rows_gemm_stats; in the reference these
// are the 1x1 Conv2d + BatchNorm2d + ReLU layers of SharedMLP on (B, C, npoint, nsample) tensors and their autograd,
// pointnet2/pytorch_utils.py:11-36, pointnet2_modules.py:251-267) for the shapes that carry the time: R = 10^5 - 10^6
// rows, K <= 160, so every operand row is touched once and the kernel should run at HBM speed (4 (K + N) bytes per row).
// The warp-level kernel reaches 1.8 - 3.2 TB/s of 6.4 (one or two CTAs per SM alike: fragment traffic through the LSU and
// 32-byte sector stores, not latency, bound it); here nothing but the fix-up pass touches the data with threads:
//
//   warp 0      one lane: TMA producer.  X tiles of 128 rows arrive as 32-column boxes (SWIZZLE_128B, K-major: the
//               layout tcgen05.mma reads) in a ring that runs across tile boundaries
//   warp 1      MMA issue, whole warp converged (umma::mma4_tf32_ss_w): per box 4 K steps against the weight slice that
//               stays in shared memory (packed once per CTA, tf32), accumulators double-buffered in tensor memory
//   warps 2-9   fix-up of each landed box in place: f(x), cvt.rna.tf32 (the tensor core would truncate)
//   warps 10-17 epilogue: TMEM -> registers -> the row's slot of a padded shared-memory tile (thread = row, two warps per
//               TMEM lane quadrant) -> ONE TMA bulk store per row (coalesced, asynchronous); column statistics read down
//               the tile's columns (thread = column x row part, running sums in registers over all tiles)
// Rows past the end of the matrix arrive as zeros, are excluded from the statistics and never stored.
#include "umma.cuh"
#include "tensor_map.cuh"
#include "rows_gemm_tc.h"

namespace eda {
namespace {

constexpr int kRows = 128;
constexpr int kBoxCols = 32;
constexpr int kBoxBytes = kRows * kBoxCols * 4;  // 16 KB
constexpr int kMaxK = 288;                       // 9 boxes
constexpr int kMaxNc = 160;                      // output columns per CTA (grid.y slices wider layers): a multiple of 32;
                                                 // columns >= N of the last slice are zero weights, clipped by the store
constexpr int kMaxRing = 8;
constexpr int kFixWarps = 8, kEpiWarps = 8;
constexpr int kThreads = (2 + kFixWarps + kEpiWarps) * 32;  // 576

struct RgTcParams {
  alignas(64) CUtensorMap map_x;
  alignas(64) CUtensorMap map_y;  // (rows, N) output, 128-row x 32-column boxes, SWIZZLE_128B: the epilogue's stores
  const float *in_scale, *in_shift, *w;
  float *y;
  double *stats;
  long long rows, w_sn, w_sk;
  int ldy, K, N, Nc, ntiles, nbox, nring;
  int tile_bufs;                                     // output tiles in shared memory (2 when they fit: the bulk stores of
                                                     // one tile are still reading it while the next one is written)
  uint32_t w_bytes, tile_off, tile_bytes, ring_off;  // shared-memory carve-up (bytes from the 1024-aligned base)
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 2-D tiled TMA store shared -> global (bulk-group completion); rows / columns outside the tensor are not written
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool kPrologue, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
rows_gemm_tc_kernel(const __grid_constant__ RgTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char *ring = base + p.ring_off;                        // [nring] boxes of 16 KB (1024-aligned)
  float *sW = reinterpret_cast<float *>(base);                    // [K / 4][Nc] float4: K-major core matrices, chunk-major
  float *tile = reinterpret_cast<float *>(base + p.tile_off);     // [tile_bufs][Nc / 32][128][32], swizzled (see the epilogue)
  __shared__ __align__(8) uint64_t full[kMaxRing], ready[kMaxRing], empty[kMaxRing], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_sc[kMaxK], s_sh[kMaxK];
  __shared__ float s_part[16][2][16];  // viewed as [nparts][2][Nc] with nparts * Nc = 256 (see the epilogue)
  static_assert(kMaxNc % 32 == 0 && kMaxNc <= 256, "one MMA per accumulator");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();
  const int K = p.K, Nc = p.Nc, nbox = p.nbox, nring = p.nring;
  const int n0 = (int)blockIdx.y * Nc;

  const uint32_t acc_stride = Nc > 128 ? 256u : 128u;  // columns between the two accumulators
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 2u * acc_stride);
  if (tid == 32) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], kFixWarps);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);
    }
    mbar_fence_init_cluster();
  }
  // weight slice -> shared memory, tf32, the chunk-major K-major layout of the forward GEMM's packed weights:
  // element (n, k) at float ((k / 4) * Nc + n) * 4 + k % 4
  for (int e = tid; e < K * Nc; e += kThreads) {
    const int k = e / Nc, n = e - k * Nc;  // consecutive threads: consecutive n (coalesced for a (K, N)-strided view)
    sW[((k >> 2) * Nc + n) * 4 + (k & 3)] =
        n0 + n < p.N ? to_tf32(__ldg(p.w + (long long)(n0 + n) * p.w_sn + (long long)k * p.w_sk)) : 0.f;
  }
  for (int k = tid; k < kMaxK; k += kThreads) {
    s_sc[k] = (kPrologue && k < K) ? __ldg(p.in_scale + k) : 0.f;
    s_sh[k] = (kPrologue && k < K) ? __ldg(p.in_shift + k) : 0.f;
  }
  umma::fence_proxy_async_smem();  // the weights are read by the tensor core (async proxy)
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;

  if (warp_u == 0) {
    // ---------------- TMA producer ----------------------------------------------------------------------------------
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 0;
      for (int t = (int)blockIdx.x; t < p.ntiles; t += (int)gridDim.x) {
        for (int j = 0; j < nbox; ++j, slot = (slot + 1 == nring) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
          mbar_wait(&empty[slot], par ^ 1u);
          mbar_arrive_expect_tx(&full[slot], (uint32_t)kBoxBytes);
          tma_load_2d(ring + (size_t)slot * kBoxBytes, &p.map_x, j * kBoxCols, t * kRows, &full[slot]);
        }
      }
    }
  } else if (warp_u == 1) {
    // ---------------- MMA issue (whole warp, converged) ----------------------------------------------------------------------
    const uint32_t idesc = umma::idesc_tf32(kRows, Nc);
    const uint64_t ad0 = umma::smem_desc_swizzled(smem_u32(ring), 16u, 1024u, 2u);
    const uint64_t bd0 = umma::smem_desc_kmajor_noswizzle(smem_u32(sW), (uint32_t)Nc * 16u, 128u);
    const uint32_t a_hi = umma::desc_hi(ad0), b_hi = umma::desc_hi(bd0);
    const uint32_t a_step = 32u >> 4, b_step = (2u * (uint32_t)Nc * 16u) >> 4, box_step = (uint32_t)kBoxBytes >> 4;
    int slot = 0;
    uint32_t par = 0;
    int it = 0;
    for (int t = (int)blockIdx.x; t < p.ntiles; t += (int)gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
      umma::fence_after_thread_sync();
      const uint32_t d = tbase + (uint32_t)buf * acc_stride;
      for (int j = 0; j < nbox; ++j, slot = (slot + 1 == nring) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
        mbar_wait(&ready[slot], par);
        umma::fence_after_thread_sync();
        const uint32_t a_lo = umma::desc_lo(ad0) + (uint32_t)slot * box_step;
        const uint32_t b_lo = umma::desc_lo(bd0) + (uint32_t)j * 4u * b_step;
        const int nks = min(kBoxCols, K - j * kBoxCols) >> 3;
        if (nks == 4) {
          umma::mma4_tf32_ss_w(d, a_lo, a_hi, a_step, b_lo, b_hi, b_step, idesc, j > 0 ? 1u : 0u);
        } else {
          for (int ks = 0; ks < nks; ++ks)
            umma::mma_tf32_ss_w(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)ks * a_step),
                                ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)ks * b_step), idesc, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma::mma_commit_w(&empty[slot]);
      }
      umma::mma_commit_w(&acc_full[buf]);
    }
  } else if (warp_u < 2 + kFixWarps) {
    // ---------------- fix-up warps: f(x) and tf32 rounding of each landed box, in place -----------------------------------------
    // thread ft owns the 16-byte chunks at position ft % 8 of rows ft / 8 + 32 i: always logical chunk (ft % 8) ^ (row & 7),
    // i.e. the same four K columns of a box
    const int ft = tid - 64;
    const int cpos = ft & 7, r0 = ft >> 3;
    const int kk = ((cpos ^ (r0 & 7)) << 2);  // first of this thread's four columns inside a box
    int slot = 0;
    uint32_t par = 0;
    for (int t = (int)blockIdx.x; t < p.ntiles; t += (int)gridDim.x) {
      for (int j = 0; j < nbox; ++j, slot = (slot + 1 == nring) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kPrologue) {
          sc = *reinterpret_cast<const float4 *>(s_sc + j * kBoxCols + kk);
          sh = *reinterpret_cast<const float4 *>(s_sh + j * kBoxCols + kk);
        }
        mbar_wait(&full[slot], par);
        float4 *bx = reinterpret_cast<float4 *>(ring + (size_t)slot * kBoxBytes) + ft;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = bx[i * 256];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (kPrologue) {
            v[i].x = fmaxf(fmaf(v[i].x, sc.x, sh.x), 0.f);
            v[i].y = fmaxf(fmaf(v[i].y, sc.y, sh.y), 0.f);
            v[i].z = fmaxf(fmaf(v[i].z, sc.z, sh.z), 0.f);
            v[i].w = fmaxf(fmaf(v[i].w, sc.w, sh.w), 0.f);
          }
          bx[i * 256] = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
        }
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(&ready[slot]);
      }
    }
  } else {
    // ---------------- epilogue warps ----------------------------------------------------------------------------------------------------
    // The output tile lives in shared memory as Nc / 32 boxes of [128 rows][32 columns], 128-byte rows with the 16-byte
    // chunks XOR-swizzled by (row & 7) — the image a SWIZZLE_128B tensor map stores from.  Phase 1, thread = (row, column
    // half): TMEM -> the row's chunks (conflict-free: 8 consecutive rows hit 8 different chunk positions).  Phase 2: ONE
    // elected thread issues one 2-D TMA store per box (rows past the end of the matrix are clipped by the tensor map; 128
    // per-row bulk stores per tile kept the TMA unit busy for ~4k cycles per tile whatever the shape), and everybody
    // takes the column statistics with thread = (column, row part): conflict-free reads down the tile's columns, running
    // sums in two registers over ALL tiles of this CTA.
    const int et = tid - (2 + kFixWarps) * 32;        // 0..255
    const int quad = warp & 3;                        // TMEM lanes 32 quad .. (hardware: warp w reads quadrant w % 4)
    const int half = et >> 7;                         // which half of the 16-column chunks
    const int r = quad * 32 + lane;
    const int nch = Nc >> 4;
    const int ch_lo = half == 0 ? 0 : (nch + 1) / 2, ch_hi = half == 0 ? (nch + 1) / 2 : nch;
    // statistics ownership: column sc, rows [sr0, sr0 + srn) of every tile (Nc in {32, 64, 128})
    const int nparts = 256 / Nc, sc = et % Nc, srn = kRows / nparts, sr0 = (et / Nc) * srn;
    const bool stat_thread = kStats && (256 % Nc == 0);
    const uint32_t s_box = (uint32_t)(sc >> 5) * (uint32_t)kBoxBytes, s_c16 = (uint32_t)((sc & 31) >> 2), s_e = (uint32_t)(sc & 3) * 4u;
    float csum = 0.f, csq = 0.f;
    int it = 0;
    for (int t = (int)blockIdx.x; t < p.ntiles; t += (int)gridDim.x, ++it) {
      const int buf = it & 1;
      unsigned char *tl = reinterpret_cast<unsigned char *>(tile) + (size_t)(p.tile_bufs == 2 ? buf : 0) * p.tile_bytes;
      mbar_wait(&acc_full[buf], (uint32_t)((it >> 1) & 1));
      umma::fence_after_thread_sync();
      if (it > 0) {
        // the stores that last read this tile buffer (issued by thread 0 of the group) have read it ...
        if (et == 0) { if (p.tile_bufs == 2) bulk_wait_read1(); else bulk_wait_read0(); }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // ... and everybody is done with that tile's columns
      }
      const uint32_t tacc = umma::tmem_addr(tbase, (uint32_t)(quad * 32), (uint32_t)buf * acc_stride);
      for (int ch = ch_lo; ch < ch_hi; ++ch) {
        const int c0 = ch << 4;
        uint32_t u[16];
        umma::tmem_ld16(tacc + (uint32_t)c0, u);
        umma::tmem_ld_wait();
        unsigned char *rowp = tl + (size_t)(c0 >> 5) * kBoxBytes + (size_t)r * 128;
        const int c16 = (c0 & 31) >> 2;  // 0 or 4: first 16-byte chunk of these 16 columns inside the box row
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<float4 *>(rowp + (((c16 + q4) ^ (r & 7)) << 4)) =
              make_float4(__uint_as_float(u[q4 * 4 + 0]), __uint_as_float(u[q4 * 4 + 1]), __uint_as_float(u[q4 * 4 + 2]),
                          __uint_as_float(u[q4 * 4 + 3]));
      }
      // the accumulator has been read: hand it back to the MMA warp
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&acc_empty[buf]);
      umma::fence_proxy_async_smem();  // the TMA store reads the tile through the async proxy
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the tile is complete
      if (et == 0) {
        for (int b = 0; b < (Nc >> 5); ++b) tma_store_2d(&p.map_y, n0 + 32 * b, t * kRows, tl + (size_t)b * kBoxBytes);
        bulk_commit_group();
      }
      if (stat_thread) {
        const long long left = p.rows - ((long long)t * kRows + sr0);  // rows of this part that exist
        const int nr = left >= srn ? srn : (left > 0 ? (int)left : 0);
        const unsigned char *col = tl + s_box + s_e;
#pragma unroll 8
        for (int i = 0; i < nr; ++i) {
          const uint32_t row_i = (uint32_t)(sr0 + i);
          const float z = *reinterpret_cast<const float *>(col + row_i * 128u + ((s_c16 ^ (row_i & 7u)) << 4));
          csum += z;
          csq = fmaf(z, z, csq);
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores have been performed
    if (kStats) {
      if (stat_thread) {
        float *sp = &s_part[0][0][0];  // [part][2][Nc]
        sp[((et / Nc) * 2 + 0) * Nc + sc] = csum;
        sp[((et / Nc) * 2 + 1) * Nc + sc] = csq;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et < Nc && n0 + et < p.N) {
        double sm = 0.0, sq = 0.0;
        const float *sp = &s_part[0][0][0];
        for (int q = 0; q < nparts; ++q) {
          sm += (double)sp[(q * 2 + 0) * Nc + et];
          sq += (double)sp[(q * 2 + 1) * Nc + et];
        }
        atomicAdd(p.stats + n0 + et, sm);
        atomicAdd(p.stats + p.N + n0 + et, sq);
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 2u * acc_stride);
}

}  // namespace

bool rows_gemm_tc_eligible(const float *x, int ldx, long long rows, int K, int N, const float *y, int ldy) {
  static const bool off = [] { const char *e = getenv("EDA_ROWS_GEMM_TC"); return e && e[0] == '0'; }();
  if (off || !encode_tiled_fn()) return false;
  if ((K & 7) || K > kMaxK || (N & 3)) return false;
  if ((ldx & 3) || (ldy & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  if (rows > 0x7fffffffLL - kRows) return false;
  return rows >= 16384;  // below that the warp-level kernel's shorter set-up wins
}

int rows_gemm_tc_launch(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                        long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y, int ldy,
                        double *stats, cudaStream_t stream) {
  RgTcParams p = {};
  if (!make_tensor_map_rows32(&p.map_x, x, rows, K, ldx, kRows, false) ||
      !make_tensor_map_rows32(&p.map_y, y, rows, N, ldy, kRows, false))
    return kRowsGemmTcDeclined;
  p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.y = y; p.stats = stats; p.rows = rows;
  p.w_sn = w_stride_n; p.w_sk = w_stride_k; p.ldy = ldy; p.K = K; p.N = N;
  p.ntiles = (int)((rows + kRows - 1) / kRows);
  p.nbox = (K + kBoxCols - 1) / kBoxCols;
  // output columns per CTA (whole 32-column boxes): the fewest slices, then the narrowest slice that covers N with them,
  // as long as a ring of >= 3 boxes fits next to the weight slice and the output tile (every slice re-reads X: K = 256,
  // N = 128 runs as two slices of 64).  With batch statistics the slice must divide 256 (thread = column x row part).
  const size_t budget = 218 * 1024;  // next to ~9 KB of static shared memory and the alignment slack
  long long nring = 0;
  const int n32 = (N + 31) & ~31;
  for (int slices = (n32 + kMaxNc - 1) / kMaxNc; slices <= 8; ++slices) {
    int nc = (((n32 / 32) + slices - 1) / slices) * 32;  // boxes per slice, rounded up
    if (stats) { int pw = 32; while (pw < nc) pw <<= 1; nc = pw; }
    if (nc > kMaxNc || (stats && nc > 128)) continue;
    p.Nc = nc;
    p.w_bytes = (uint32_t)(((size_t)K * nc * 4 + 1023) & ~(size_t)1023);
    p.tile_bytes = (uint32_t)((nc / 32) * kBoxBytes);
    p.tile_off = p.w_bytes;
    p.tile_bufs = 1;
    p.ring_off = p.w_bytes + p.tile_bytes;
    nring = ((long long)budget - 1024 - (long long)p.ring_off) / kBoxBytes;
    // a second output tile when it leaves a ring of one tile's boxes + 2 (at least 4)
    const long long want = p.nbox + 2 > 4 ? p.nbox + 2 : 4;
    if (((long long)budget - 1024 - (long long)p.ring_off - (long long)p.tile_bytes) / kBoxBytes >= want) {
      p.tile_bufs = 2;
      p.ring_off += p.tile_bytes;
      nring = ((long long)budget - 1024 - (long long)p.ring_off) / kBoxBytes;
    }
    if (nring >= 3) break;
  }
  if (nring > kMaxRing) nring = kMaxRing;
  if (nring < 3) return kRowsGemmTcDeclined;
  if (stats && (256 % p.Nc)) return kRowsGemmTcDeclined;  // the statistics pass deals 256 threads as columns x row parts
  p.nring = (int)nring;
  const size_t smem = (size_t)p.ring_off + (size_t)p.nring * kBoxBytes + 1024;
  const int ny = (N + p.Nc - 1) / p.Nc;
  long long gx = sm_count() / ny;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  const dim3 grid((unsigned)gx, (unsigned)ny);
  static SmemAttr attr[4];
  const int which = (in_scale ? 1 : 0) | (stats ? 2 : 0);
  switch (which) {
    case 0:
      EDA_CUDA_TRY(attr[0].ensure(rows_gemm_tc_kernel<false, false>, smem), "rows_gemm_tc smem attr");
      rows_gemm_tc_kernel<false, false><<<grid, kThreads, smem, stream>>>(p);
      break;
    case 1:
      EDA_CUDA_TRY(attr[1].ensure(rows_gemm_tc_kernel<true, false>, smem), "rows_gemm_tc smem attr");
      rows_gemm_tc_kernel<true, false><<<grid, kThreads, smem, stream>>>(p);
      break;
    case 2:
      EDA_CUDA_TRY(attr[2].ensure(rows_gemm_tc_kernel<false, true>, smem), "rows_gemm_tc smem attr");
      rows_gemm_tc_kernel<false, true><<<grid, kThreads, smem, stream>>>(p);
      break;
    default:
      EDA_CUDA_TRY(attr[3].ensure(rows_gemm_tc_kernel<true, true>, smem), "rows_gemm_tc smem attr");
      rows_gemm_tc_kernel<true, true><<<grid, kThreads, smem, stream>>>(p);
      break;
  }
  return check_launch("rows_gemm_tc_kernel");
}

}  // namespace eda
