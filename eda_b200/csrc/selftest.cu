// Hardware self-test of the tcgen05 building blocks in umma.cuh (descriptor encodings, the
// chunk-major K-major shared-memory layout, TMEM lane/column addressing, A-from-TMEM).  One CTA
// computes D[128 x N] = A[128 x K] * W[N x K]^T with kind::tf32.  Exposed through the C ABI so that
// tests/test_gpu_umma.py can check it against a CPU tf32-rounded reference before any fused
// kernel relies on these conventions.
#include "umma.cuh"

namespace eda {
namespace {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ W, int N, int K, int mode,
                     float *__restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *sA = reinterpret_cast<float4 *>(smem_raw);  // [K/4][128]
  float4 *sW = sA + (K / 4) * 128;                    // [K/4][N]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init_cluster();
  }
  if (mode == 2) {
    // MN-major B: float4 [N/4][K], one unit = 4 consecutive n of one k
    for (int i = tid; i < (N / 4) * K; i += 128) {
      const int j = i / K, k = i % K;
      sW[i] = make_float4(W[(size_t)(4 * j) * K + k], W[(size_t)(4 * j + 1) * K + k], W[(size_t)(4 * j + 2) * K + k],
                          W[(size_t)(4 * j + 3) * K + k]);
    }
  } else {
    for (int i = tid; i < (K / 4) * N; i += 128) {
      const int c = i / N, n = i % N;
      sW[i] = *reinterpret_cast<const float4 *>(W + (size_t)n * K + c * 4);
    }
  }
  if (mode == 0 || mode == 2) {
    for (int c = 0; c < K / 4; ++c) sA[c * 128 + tid] = *reinterpret_cast<const float4 *>(A + (size_t)tid * K + c * 4);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32);
  constexpr uint32_t kAcol = 256;  // A operand lives at columns [256, 256+K) in mode 1
  if (mode == 1) {
    for (int c0 = 0; c0 < K; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(A[(size_t)tid * K + c0 + j]);
      umma::tmem_st16(umma::tmem_addr(tbase, lane_base, kAcol + c0), v);
    }
    umma::tmem_st_wait();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = mode == 2 ? umma::idesc_tf32_bmn(128, N) : umma::idesc_tf32(128, N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t bdesc = mode == 2
                                 ? umma::smem_desc_kmajor_noswizzle(smem_u32(sW + ks * 8), 128u, (uint32_t)K * 16u)
                                 : umma::smem_desc_kmajor_noswizzle(smem_u32(sW + ks * 2 * N), (uint32_t)N * 16u, 128u);
      if (mode == 0 || mode == 2) {
        const uint64_t adesc = umma::smem_desc_kmajor_noswizzle(smem_u32(sA + ks * 2 * 128), 128u * 16u, 128u);
        umma::mma_tf32_ss(tbase, adesc, bdesc, idesc, ks > 0);
      } else {
        umma::mma_tf32_ts(tbase, tbase + kAcol + ks * 8, bdesc, idesc, ks > 0);
      }
    }
    umma::mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    umma::tmem_ld16(umma::tmem_addr(tbase, lane_base, c0), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Layout probe: one MMA (K = 8) with A = [I_8; 0] and the B region of shared memory filled with
// float(i) at float index i, so D[k][n] = the shared-memory float index the hardware reads as B(n, k)
// for the given descriptor strides / majorness.
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(int N, int lbo, int sbo, int b_mn, int layout, int start_off, float *__restrict__ D) {
  __shared__ __align__(128) float4 sA[2 * 128];
  __shared__ __align__(1024) float sB[4096];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init_cluster();
  }
  for (int c = 0; c < 2; ++c) {
    float f[4];
    for (int e = 0; e < 4; ++e) f[e] = (tid == c * 4 + e) ? 1.f : 0.f;
    sA[c * 128 + tid] = make_float4(f[0], f[1], f[2], f[3]);
  }
  for (int i = tid; i < 4096; i += 128) sB[i] = (float)(i & 2047);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_slot;
  if (tid == 0) {
    const uint32_t idesc = b_mn ? umma::idesc_tf32_bmn(128, N) : umma::idesc_tf32(128, N);
    const uint64_t adesc = umma::smem_desc_kmajor_noswizzle(smem_u32(sA), 128u * 16u, 128u);
    const uint64_t bdesc = umma::smem_desc_swizzled(smem_u32(sB) + (uint32_t)start_off, (uint32_t)lbo, (uint32_t)sbo, (uint32_t)layout);
    umma::mma_tf32_ss(tbase, adesc, bdesc, idesc, 0);
    umma::mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_thread_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    umma::tmem_ld16(umma::tmem_addr(tbase, (uint32_t)(warp * 32), c0), v);
    umma::tmem_ld_wait();
    if (tid < 8)
      for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}


// MMA issue / execution rate probe (measurement aid): one CTA per block of the grid issues `iters` kind::tf32 MMAs of
// shape 128 x N x 8 back to back from one thread (fixed operand addresses, accumulate on) and reports the cycles
// between the first issue and the completion of the last one (clock64).  a_mode 0: A from shared memory (K-major,
// no swizzle), 1: A from tensor memory, 2: A from shared memory in the SWIZZLE_128B layout the linear kernel stages.
// b_swz != 0: B in SWIZZLE_128B too (rows of 128 bytes), else the chunk-major no-swizzle layout of the packed weights.
// precomputed != 0: descriptors built once outside the loop (isolates the instruction overhead of the issue loop).
__global__ void __launch_bounds__(1024, 1)
umma_rate_kernel(int N, int a_mode, int b_swz, int precomputed_arg, int iters, long long *__restrict__ cycles) {
  // precomputed_arg: bits 0-7 the mode below; bits 8-19 a second accumulator's column offset (0 = one accumulator):
  // MMAs then go in pairs sharing A, alternating accumulators and B halves, as the linear kernel does for N > 256;
  // bits 20-30 the row count of the B tile the MMA's N rows are cut from (LBO = rows * 16 bytes; 0 = N)
  const int precomputed = precomputed_arg & 127;
  const uint32_t acc_off = (uint32_t)(precomputed_arg >> 8) & 4095u;
  const uint32_t b_rows = ((uint32_t)precomputed_arg >> 20) ? ((uint32_t)precomputed_arg >> 20) : (uint32_t)N;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar, spin_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t adesc_tab[16], bdesc_tab[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&spin_bar, 1);
    mbar_fence_init_cluster();
  }
  float *f = reinterpret_cast<float *>(base);
  // 128 KB of operands: zeros, or (mode bit 7) pseudo-random tf32 values in (-1, 1) — real data toggles the datapath
  for (int i = tid; i < 32768; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    f[i] = (precomputed_arg & 128) ? __uint_as_float(__float_as_uint((float)(int)(h & 0xffffu) * (1.f / 32768.f) - 1.f) & 0xffffe000u) : 0.f;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_slot;
  if (precomputed == 4) {
    // warp-converged issue loop (umma::mma_tf32_ss_w): all of warp 0 runs it, descriptors by uniform arithmetic, fresh
    // operand addresses as in mode 2
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // tells the compiler the branch below is warp-uniform
    if (warp_u == 0) {
      const uint32_t idesc = umma::idesc_tf32(128, N);
      const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base) + 65536u;
      const uint32_t lbo_w = b_rows * 16u;
      const uint32_t b_steps = 65536u / (2u * lbo_w);
      const uint64_t a0 = umma::smem_desc_swizzled(a_addr, 16u, 1024u, 2u);
      const uint64_t b0 = umma::smem_desc_kmajor_noswizzle(b_addr, lbo_w, 128u);
      const uint64_t b_step = (uint64_t)((2u * lbo_w) >> 4);
      const long long t0 = clock64();
      uint32_t ib = 0;
      for (int i = 0; i < iters; ++i) {
        const uint32_t ia = (uint32_t)i & 15u;
        const uint64_t ad = a0 + (uint64_t)((ia >> 2) * 1024u + (ia & 3u) * 2u);
        const uint64_t bd = b0 + (uint64_t)ib * b_step;
        if (a_mode == 1) umma::mma_tf32_ts_w(tbase + (uint32_t)(i & 1) * acc_off, tbase + 288u + (ia & 3u) * 8u, bd, idesc, 1u);
        else umma::mma_tf32_ss_w(tbase + (uint32_t)(i & 1) * acc_off, ad, bd, idesc, 1u);
        ib = ib + 1 == b_steps ? 0u : ib + 1;
      }
      const long long t1 = clock64();
      umma::mma_commit_w(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 0) {
        cycles[2 * blockIdx.x] = t1 - t0;
        cycles[2 * blockIdx.x + 1] = t2 - t0;
        mbar_arrive_expect_tx(&spin_bar, 0);
      }
    } else if (warp_u >= 4) {
      mbar_wait(&spin_bar, 0);
    }
  } else
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(128, N);
    // precomputed == 2: every MMA reads FRESH shared-memory addresses (16 distinct A K-steps over 64 KB, as many B
    // K-steps as fit into 64 KB) — what a real K loop does; the other modes cycle through 4 K-steps of one tile
    const bool fresh = precomputed == 2 || precomputed == 3;
    const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base) + (fresh ? 65536u : 16384u);
    const uint32_t lbo_w = b_rows * 16u;
    const uint32_t b_swz_tile = ((uint32_t)N * 128u + 1023u) & ~1023u;
    const uint32_t b_steps = !fresh ? 4u : b_swz ? 4u * (65536u / b_swz_tile) : 65536u / (2u * lbo_w);
    const uint64_t a0 = a_mode == 2 ? umma::smem_desc_swizzled(a_addr, 16u, 1024u, 2u)
                                    : umma::smem_desc_kmajor_noswizzle(a_addr, 128u * 16u, 128u);
    const uint64_t b0 = b_swz ? umma::smem_desc_swizzled(b_addr, 16u, 1024u, 2u)
                              : umma::smem_desc_kmajor_noswizzle(b_addr, lbo_w, 128u);
    if (fresh) {
      // descriptor tables built outside the timed loop: the loop itself is two shared-memory loads and the MMA
      for (uint32_t i = 0; i < 16u; ++i) {
        const uint32_t ia = acc_off ? (i >> 1) : i, ib = (acc_off ? (i >> 1) : i) % b_steps;
        const uint32_t half = acc_off ? (i & 1u) * (uint32_t)N * 16u : 0u;
        adesc_tab[i] = a_mode == 2 ? umma::smem_desc_swizzled(a_addr + (ia >> 2) * 16384u + (ia & 3u) * 32u, 16u, 1024u, 2u)
                                   : umma::smem_desc_kmajor_noswizzle(a_addr + ia * 2u * 2048u, 128u * 16u, 128u);
        bdesc_tab[i] = b_swz ? umma::smem_desc_swizzled(b_addr + (ib >> 2) * (((uint32_t)N * 128u + 1023u) & ~1023u) + (ib & 3u) * 32u, 16u, 1024u, 2u)
                             : umma::smem_desc_kmajor_noswizzle(b_addr + ib * 2u * lbo_w + half, lbo_w, 128u);
      }
      const long long t0 = clock64();
      if (precomputed == 3 && a_mode != 1) {
        // same fresh addresses, but the descriptors of four MMAs are fetched into distinct registers before the four
        // issues: separates "new operand addresses are slow" from "rewriting the descriptor registers of an MMA that
        // is still queued stalls the issuing thread"
        for (int i = 0; i < iters; i += 4) {
          const int b = i & 12;
          const uint64_t a0_ = adesc_tab[b], a1_ = adesc_tab[b + 1], a2_ = adesc_tab[b + 2], a3_ = adesc_tab[b + 3];
          const uint64_t b0_ = bdesc_tab[b], b1_ = bdesc_tab[b + 1], b2_ = bdesc_tab[b + 2], b3_ = bdesc_tab[b + 3];
          umma::mma_tf32_ss(tbase, a0_, b0_, idesc, 1u);
          umma::mma_tf32_ss(tbase + acc_off, a1_, b1_, idesc, 1u);
          umma::mma_tf32_ss(tbase, a2_, b2_, idesc, 1u);
          umma::mma_tf32_ss(tbase + acc_off, a3_, b3_, idesc, 1u);
        }
      } else
#pragma unroll 4
      for (int i = 0; i < iters; ++i) {
        const uint64_t ad = adesc_tab[i & 15], bd = bdesc_tab[i & 15];
        const uint32_t d = tbase + (uint32_t)(i & 1) * acc_off;
        if (a_mode == 1) umma::mma_tf32_ts(d, tbase + 288u + (uint32_t)(i & 3) * 8u, bd, idesc, 1u);
        else umma::mma_tf32_ss(d, ad, bd, idesc, 1u);
      }
      const long long t1 = clock64();
      umma::mma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      cycles[2 * blockIdx.x] = t1 - t0;
      cycles[2 * blockIdx.x + 1] = t2 - t0;
      mbar_arrive_expect_tx(&spin_bar, 0);
    } else {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t ks = (uint32_t)(i & 3);
      uint64_t ad = a0, bd = b0;
      if (!precomputed) {  // what the production loops do: rebuild both descriptors per K step
        ad = a_mode == 2 ? umma::smem_desc_swizzled(a_addr + ks * 32u, 16u, 1024u, 2u)
                         : umma::smem_desc_kmajor_noswizzle(a_addr + ks * 2u * 2048u, 128u * 16u, 128u);
        bd = b_swz ? umma::smem_desc_swizzled(b_addr + ks * 32u, 16u, 1024u, 2u)
                   : umma::smem_desc_kmajor_noswizzle(b_addr + ks * 2u * lbo_w, lbo_w, 128u);
      }
      if (a_mode == 1) umma::mma_tf32_ts(tbase, tbase + 288u + ks * 8u, bd, idesc, 1u);
      else umma::mma_tf32_ss(tbase, ad, bd, idesc, 1u);
    }
    const long long t1 = clock64();
    umma::mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    cycles[2 * blockIdx.x] = t1 - t0;      // issue loop
    cycles[2 * blockIdx.x + 1] = t2 - t0;  // until the last MMA has completed
    mbar_arrive_expect_tx(&spin_bar, 0);   // releases the spinner warps
    }
  } else if (warp >= 4) {
    // extra warps (blockDim > 128): wait on a shared-memory barrier for the whole measurement, like the staging /
    // producer warps of the GEMM kernels do while the issuer thread works
    mbar_wait(&spin_bar, 0);
  }
  __syncthreads();
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

}  // namespace
}  // namespace eda

extern "C" int eda_selftest_umma_probe(int N, int lbo_bytes, int sbo_bytes, int b_mn_major, int layout_type,
                                       int start_offset_bytes, float *D, void *stream) {
  using namespace eda;
  if (!D || N < 16 || N > 256 || N % 16) return EDA_ERR_INVALID_ARGUMENT;
  umma_probe_kernel<<<1, 128, 0, as_stream(stream)>>>(N, lbo_bytes, sbo_bytes, b_mn_major, layout_type, start_offset_bytes, D);
  return check_launch("umma_probe_kernel");
}

extern "C" int eda_selftest_umma(const float *A, const float *W, int N, int K, int mode, float *D, void *stream) {
  using namespace eda;
  if (!A || !W || !D) return EDA_ERR_INVALID_ARGUMENT;
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16 || (mode != 0 && mode != 1))
    return EDA_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(K / 4) * (128 + N) * sizeof(float4);
  EDA_CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "selftest smem attr");
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(A, W, N, K, mode, D);
  return check_launch("umma_selftest_kernel");
}

extern "C" int eda_selftest_umma_rate(int N, int a_mode, int b_swizzled, int precomputed, int iters, int ctas,
                                      int waiting_warps, long long *cycles_device, void *stream) {
  using namespace eda;
  if (!cycles_device || N < 16 || N > 256 || N % 16 || iters < 1 || ctas < 1 || a_mode < 0 || a_mode > 2 ||
      waiting_warps < 0 || waiting_warps > 28)
    return EDA_ERR_INVALID_ARGUMENT;
  const size_t smem = 131072 + 1024;
  EDA_CUDA_TRY(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "umma rate smem attr");
  umma_rate_kernel<<<ctas, 128 + 32 * waiting_warps, smem, as_stream(stream)>>>(N, a_mode, b_swizzled, precomputed, iters,
                                                                               cycles_device);
  return check_launch("umma_rate_kernel");
}
