// Weight gradient on the tcgen05 tensor cores (sm_100a):  dW[N x K] += dY[R x N]^T X[R x K],  db[N] += column sums of dY.
//
// Replaces, for the large row counts of the attention layers' linears (8192 visual tokens, 2048 queries per step), the
// warp-level mma.sync kernel of grad_ops.cu behind the same entry point (eda_wgrad); in the reference this is autograd's
// mm / addmm backward of every nn.Linear (models/encoder_decoder_layers.py:47-71,298-328): one cuBLAS GEMM plus one
// column reduction per weight.
//
// The contraction index is the ROW of both row-major operands, so both are "MN-major" for the tensor core.  For 32-bit
// operands the tensor core reads MN-major tiles only in the "128-byte swizzle with 32-byte atoms" layout (UMMA layout
// type 1; types 0 / 2 / 4 / 6 return zeros — scripts/probe_umma_mn.py): one 128-byte row per contraction index
// holding 32 consecutive features, its four 32-byte units XOR-swizzled by (row & 3), four rows = one 512-byte group, the
// next 32 features one box further.  That is exactly what a 2-D TMA box of 32 rows x 32 columns with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B puts into shared memory.  No transposition, no per-thread staging: dY and X tiles
// go global -> shared -> tensor core as they lie in memory.
//
// CTA = one tile of 128 output rows (features n of dY) x one K chunk (<= 480 features k of X) x one slice of the R rows:
//   warp 0     one lane: TMA producer, ring of 32-row stages: 4 boxes of dY + ceil(kw / 32) boxes of X per stage
//   warp 1     MMA issue, whole warp converged (umma::mma4_tf32_ss_w): per stage 4 K steps of 8 rows, each one or two
//              MMAs of 128 x (<= 256) columns; fp32 accumulators in tensor memory
//   warps 2-9  round the landed operands to tf32 in place (cvt.rna; the tensor core would truncate; X optionally through
//              relu(x * scale[k] + shift[k]) first — the folded BatchNorm + ReLU of the SA layers), then the epilogue:
//              TMEM -> registers -> red.global.add (v4) into dW; thread = output row
// The bias gradient (exact fp32 column sums of dY, before rounding) is taken by the rounding warps on the way: a thread
// meets the same four features of every dY box in every stage, keeps their running sums in registers and adds them into
// a shared-memory row at the end.
// Row slices of one output tile meet in global memory through fp32 atomics (as in the mma.sync kernel): dW is
// ACCUMULATED, zero it first.
#include "umma.cuh"
#include "tensor_map.cuh"
#include "wgrad_tc.h"

namespace eda {
namespace {

constexpr int kTcRows = 32;                // contraction rows per ring stage = TMA box height
constexpr int kBoxBytes = kTcRows * 128;   // 32 rows x 32 fp32
constexpr int kTileN = 128;                // output rows (features of dY) per CTA = MMA M
constexpr int kYBoxes = kTileN / 32;
constexpr int kTcThreads = 320;
constexpr int kRoundWarps = 8;
constexpr int kTcMaxProbs = 6;
constexpr int kTcMaxStages = 4;
constexpr int kTcMaxBoxes = 16;            // X boxes per K chunk: 512 TMEM columns, two MMAs of <= 256 columns
constexpr int kTcMaxChunk = 32 * kTcMaxBoxes;
constexpr int kScratchBytes = kRoundWarps * 32 * kYBoxes * 16;  // bias partials: [warp][lane][box] float4 = 16 KB

struct TcProblem {
  const float *x_scale, *x_shift;  // optional (K): x is consumed as relu(x * x_scale[k] + x_shift[k])
  float *dw, *db;
  long long rows;
  int ldw;
  int vec;  // dw rows are 16-byte aligned: red.global.add.v4.f32
};
struct TcParams {
  alignas(64) CUtensorMap map_y[kTcMaxProbs];
  alignas(64) CUtensorMap map_x[kTcMaxProbs];
  TcProblem pr[kTcMaxProbs];
  int nprobs, N, K, ntiles, kchunks, kchunk, splits, nstages;
  uint32_t stage_bytes;
  int ep_whole;  // the epilogue tile holds all accumulator columns at once (else one MMA piece at a time)
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bulk asynchronous reduction shared -> global on the TMA engine: dst[i] += src[i] over `bytes` of fp32 (16-byte aligned
// both sides, size a multiple of 16).  The adds happen in L2 at copy granularity instead of as per-thread atomics.
__device__ __forceinline__ void bulk_reduce_add_f32(float *dst_gmem, const float *src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Development aid: phase timestamps (clock64) of CTA 0 of the most recent launch (eda_debug_timestamps_wgrad): [0] entry,
// [1] set-up done, [2] first stage landed, [3] first stage rounded, [4] accumulators complete, [5] bias sums done,
// [6] epilogue done, [7] exit.
__device__ long long g_wg_ts[16];
#define WG_TS(i, t) do { if (blockIdx.x == 0 && tid == (t)) g_wg_ts[i] = clock64(); } while (0)

__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char *ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // swizzle atoms: 1024-byte aligned
  __shared__ __align__(8) uint64_t full[kTcMaxStages], ready[kTcMaxStages], empty[kTcMaxStages], done;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_sc[kTcMaxChunk], s_sh[kTcMaxChunk];  // the prologue's scale / shift of this K chunk

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = umma::uniform_warp_index();
  WG_TS(0, 0);
  // blockIdx.x = ((problem * splits + split) * kchunks + kc) * ntiles + nt
  int bid = (int)blockIdx.x;
  const int nt = bid % p.ntiles; bid /= p.ntiles;
  const int kc = bid % p.kchunks; bid /= p.kchunks;
  const int split = bid % p.splits;
  const int pi = bid / p.splits;
  const TcProblem &pr = p.pr[pi];
  const int n0 = nt * kTileN, k0 = kc * p.kchunk;
  const int kw = min(p.kchunk, p.K - k0);   // features of X in this chunk
  const int nb = (kw + 31) >> 5;            // X boxes per stage
  const long long chunks_total = (pr.rows + kTcRows - 1) / kTcRows;
  const long long per = (chunks_total + p.splits - 1) / p.splits;
  const long long c_lo = (long long)split * per;
  const long long c_hi = c_lo + per < chunks_total ? c_lo + per : chunks_total;
  if (c_lo >= c_hi) return;  // (whole CTA; nothing allocated yet)

  // accumulator columns [0, 32 nb): one MMA if they fit 256 columns, else two pieces cut at a box boundary
  const int cols = 32 * nb;
  const int n_a = cols <= 256 ? cols : 32 * ((nb + 1) / 2);
  const int n_b = cols - n_a;
  const uint32_t tcols = cols <= 32 ? 32u : cols <= 64 ? 64u : cols <= 128 ? 128u : cols <= 256 ? 256u : 512u;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, tcols);
  if (tid == 32) {
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], kRoundWarps);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    mbar_fence_init_cluster();
  }
  pdl_launch_dependents();
  pdl_wait();  // dY / X (and the folded BatchNorm terms) come from the kernels before this one
  const bool prologue = pr.x_scale != nullptr;
  if (prologue)
    for (int i = tid; i < kTcMaxChunk; i += kTcThreads) {
      const bool in = i < kw;  // features past the chunk: scale = shift = 0, so the zero-filled columns stay zero
      s_sc[i] = in ? __ldg(pr.x_scale + k0 + i) : 0.f;
      s_sh[i] = in ? __ldg(pr.x_shift + k0 + i) : 0.f;
    }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tbase = tmem_slot;
  WG_TS(1, 0);

  if (warp_u == 0) {
    // ---------------- TMA producer ----------------------------------------------------------------------------------
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 0;
      for (long long c = c_lo; c < c_hi; ++c, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
        mbar_wait(&empty[slot], par ^ 1u);
        unsigned char *st = ring + (size_t)slot * p.stage_bytes;
        mbar_arrive_expect_tx(&full[slot], (uint32_t)((kYBoxes + nb) * kBoxBytes));  // boxes count in full, zero fill included
        const int r0 = (int)(c * kTcRows);
#pragma unroll
        for (int j = 0; j < kYBoxes; ++j) tma_load_2d(st + (size_t)j * kBoxBytes, &p.map_y[pi], n0 + 32 * j, r0, &full[slot]);
        for (int j = 0; j < nb; ++j)
          tma_load_2d(st + (size_t)(kYBoxes + j) * kBoxBytes, &p.map_x[pi], k0 + 32 * j, r0, &full[slot]);
      }
    }
  } else if (warp_u == 1) {
    // ---------------- MMA issue (whole warp, converged; see umma::mma4_tf32_ss_w) ------------------------------------------------
    constexpr uint32_t kMnMajor = (1u << 15) | (1u << 16);  // A and B both MN-major
    const uint32_t idesc_a = umma::idesc_tf32(kTileN, n_a) | kMnMajor;
    const uint32_t idesc_b = umma::idesc_tf32(kTileN, n_b > 0 ? n_b : 16) | kMnMajor;
    // MN-major, 128-byte swizzle with 32-byte atoms (layout type 1): 32 features per 128-byte row; leading offset = the
    // next 32 features (one box, 4096 bytes), stride offset = the next group of 4 contraction rows (512 bytes); a K step
    // of 8 rows advances the start address by 1024 bytes
    const uint64_t ad0 = umma::smem_desc_swizzled(smem_u32(ring), kBoxBytes, 512u, 1u);
    const uint64_t bd0 = umma::smem_desc_swizzled(smem_u32(ring) + kYBoxes * kBoxBytes, kBoxBytes, 512u, 1u);
    const uint32_t a_hi = umma::desc_hi(ad0), b_hi = umma::desc_hi(bd0);
    const uint32_t k_step = 1024u >> 4, stage_step = p.stage_bytes >> 4;
    const uint32_t b_piece = ((uint32_t)(n_a / 32) * kBoxBytes) >> 4;
    int slot = 0;
    uint32_t par = 0;
    for (long long c = c_lo; c < c_hi; ++c, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
      mbar_wait(&ready[slot], par);
      umma::fence_after_thread_sync();
      const uint32_t a_lo = umma::desc_lo(ad0) + (uint32_t)slot * stage_step;
      const uint32_t b_lo = umma::desc_lo(bd0) + (uint32_t)slot * stage_step;
      const uint32_t acc = c > c_lo ? 1u : 0u;
      umma::mma4_tf32_ss_w(tbase, a_lo, a_hi, k_step, b_lo, b_hi, k_step, idesc_a, acc);
      if (n_b > 0) umma::mma4_tf32_ss_w(tbase + (uint32_t)n_a, a_lo, a_hi, k_step, b_lo + b_piece, b_hi, k_step, idesc_b, acc);
      umma::mma_commit_w(&empty[slot]);
    }
    umma::mma_commit_w(&done);
  } else {
    // ---------------- rounding warps, then the epilogue ----------------------------------------------------------------------
    // thread rt owns, in every box of every stage, the 16-byte chunk at position rt % 8 of row rt / 8 (a box is 256
    // chunks): with the swizzle that is half (rt & 1) of logical 32-byte unit ((rt % 8) / 2) ^ (row & 3), i.e. always the
    // same four features of a box
    const int rt = tid - 64;  // 0..255
    static_assert(kRoundWarps * 32 == kBoxBytes / 16, "one chunk per thread and box");
    const bool want_db = pr.db != nullptr && kc == 0;
    float4 bsum[kYBoxes];
#pragma unroll
    for (int j = 0; j < kYBoxes; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    int slot = 0;
    uint32_t par = 0;
    for (long long c = c_lo; c < c_hi; ++c, slot = (slot + 1 == p.nstages) ? 0 : slot + 1, par ^= (slot == 0) ? 1u : 0u) {
      mbar_wait(&full[slot], par);
      if (c == c_lo) WG_TS(2, 64);
      float4 *st = reinterpret_cast<float4 *>(ring + (size_t)slot * p.stage_bytes) + rt;
#pragma unroll
      for (int j = 0; j < kYBoxes; ++j) {
        float4 v = st[j * (kBoxBytes / 16)];
        bsum[j].x += v.x; bsum[j].y += v.y; bsum[j].z += v.z; bsum[j].w += v.w;
        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
        st[j * (kBoxBytes / 16)] = v;
      }
      if (prologue) {
        // x -> relu(x * scale + shift) on the way (the folded BatchNorm + ReLU of the layer that produced x).  Rows past
        // the end arrive as zeros and become relu(shift): they meet zero rows of dY, so they add nothing
        const int feat = ((((rt & 7) >> 1) ^ ((rt >> 3) & 3)) << 3) | ((rt & 1) << 2);  // this thread's 4 features in a box
#pragma unroll 4
        for (int j = 0; j < nb; ++j) {
          const float4 sc = *reinterpret_cast<const float4 *>(s_sc + 32 * j + feat);
          const float4 sh = *reinterpret_cast<const float4 *>(s_sh + 32 * j + feat);
          float4 v = st[(kYBoxes + j) * (kBoxBytes / 16)];
          v.x = to_tf32(fmaxf(fmaf(v.x, sc.x, sh.x), 0.f)); v.y = to_tf32(fmaxf(fmaf(v.y, sc.y, sh.y), 0.f));
          v.z = to_tf32(fmaxf(fmaf(v.z, sc.z, sh.z), 0.f)); v.w = to_tf32(fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
          st[(kYBoxes + j) * (kBoxBytes / 16)] = v;
        }
      } else {
#pragma unroll 4
        for (int j = kYBoxes; j < kYBoxes + nb; ++j) {
          float4 v = st[j * (kBoxBytes / 16)];
          v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
          st[j * (kBoxBytes / 16)] = v;
        }
      }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&ready[slot]);
      if (c == c_lo) WG_TS(3, 64);
    }
    mbar_wait(&done, 0);
    umma::fence_after_thread_sync();
    __syncwarp();
    WG_TS(4, 64);
    if (want_db) {
      // all MMAs have completed: the ring is free (every rounding warp's last write precedes its last `ready` arrival,
      // which precedes the final commit onto `done` — an mbarrier chain; compute-sanitizer's racecheck, which only
      // follows bar.sync, reports these writes against the loop's as hazards).  Partials -> ring[warp][lane][box], then thread f < 128 adds up the
      // 8 warps x 4 lanes that met feature f (lane = (row & 3) * 8 + chunk position; shared-memory float atomics would
      // be 32-way contended compare-and-swap loops)
      float4 *scratch = reinterpret_cast<float4 *>(ring);
#pragma unroll
      for (int j = 0; j < kYBoxes; ++j) scratch[rt * kYBoxes + j] = bsum[j];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (rt < kTileN && n0 + rt < p.N) {
        const int j = rt >> 5, ch = (rt & 31) >> 2, e = rt & 3;
        const float *sf = reinterpret_cast<const float *>(ring);
        float acc = 0.f;
#pragma unroll
        for (int r3 = 0; r3 < 4; ++r3) {
          const int ln = r3 * 8 + ((((ch >> 1) ^ r3) << 1) | (ch & 1));
#pragma unroll
          for (int w = 0; w < kRoundWarps; ++w) acc += sf[((w * 32 + ln) * kYBoxes + j) * 4 + e];
        }
        atomicAdd(pr.db + n0 + rt, acc);
      }
    }
    WG_TS(5, 64);
    // Accumulators -> dW in two phases per MMA piece, through a padded shared-memory tile behind the bias scratch (the ring
    // is free).  Phase 1, thread = output row n (TMEM lane; warp w reads lanes 32 (w % 4).., the two warps of a quadrant
    // split the columns).  Phase 2, thread = row: ONE bulk reduce-add (cp.reduce.async.bulk .add.f32) of the row's
    // columns into dW — the TMA engine adds in L2 at line granularity.  As per-thread red.global.add.v4 the same traffic
    // ran at ~1 atomic transaction per cycle and SM, coalesced or not: 9.5k - 14k cycles of a 24k-cycle kernel.
    const int quad = warp & 3, half = (warp - 2) >> 2, w8 = warp - 2;
    const int r = quad * 32 + lane;
    const uint32_t trow = umma::tmem_addr(tbase, (uint32_t)(quad * 32), 0);
    float *tile = reinterpret_cast<float *>(ring + kScratchBytes);
    const int npieces = (p.ep_whole || n_b == 0) ? 1 : 2;
    for (int piece = 0; piece < npieces; ++piece) {
      const int pc0 = piece ? n_a : 0, pcols = npieces == 1 ? cols : (piece ? n_b : n_a);  // multiples of 32
      const int pitch = pcols + 4;                                 // = 4 (mod 32): conflict-free float4 rows
      const int nch = pcols >> 4;
      const int ch_lo = half == 0 ? 0 : nch / 2, ch_hi = half == 0 ? nch / 2 : nch;
      for (int ch = ch_lo; ch < ch_hi; ++ch) {
        uint32_t u[16];
        umma::tmem_ld16(trow + (uint32_t)(pc0 + ch * 16), u);
        umma::tmem_ld_wait();
        float4 *d = reinterpret_cast<float4 *>(tile + (size_t)r * pitch + ch * 16);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          d[q4] = make_float4(__uint_as_float(u[q4 * 4 + 0]), __uint_as_float(u[q4 * 4 + 1]), __uint_as_float(u[q4 * 4 + 2]),
                              __uint_as_float(u[q4 * 4 + 3]));
      }
      umma::fence_proxy_async_smem();  // the bulk reduction reads the tile through the async proxy
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n4 = max(0, min(pcols, kw - pc0)) >> 2;  // K is a multiple of 4: whole float4 groups inside the matrix
      if (pr.vec) {
        if (rt < kTileN && n0 + rt < p.N && n4 > 0) {
          bulk_reduce_add_f32(pr.dw + (size_t)(n0 + rt) * pr.ldw + k0 + pc0, tile + (size_t)rt * pitch, (uint32_t)n4 * 16u);
          bulk_commit_group();
          bulk_wait_group_read0();  // the tile row has been read: it may be rewritten
        }
      } else {  // dW rows not 16-byte aligned: scalar atomics, warp = row
        for (int rr = 0; rr < kTileN / kRoundWarps; ++rr) {
          const int row = w8 * (kTileN / kRoundWarps) + rr;
          const int n = n0 + row;
          if (n >= p.N) break;
          float *dst = pr.dw + (size_t)n * pr.ldw + k0 + pc0;
          const float *src = tile + (size_t)row * pitch;
          for (int c = lane; c < 4 * n4; c += 32) atomicAdd(dst + c, src[c]);
        }
      }
      if (piece + 1 < npieces) asm volatile("bar.sync 1, 256;" ::: "memory");  // the tile is rewritten by the second piece
    }
    bulk_wait_group0();  // reductions of this thread have been performed before the CTA retires
  }
  WG_TS(6, 64);
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, tcols);
  WG_TS(7, 0);
}

}  // namespace

int wgrad_tc_timestamps(long long *host_out, int n) {
  if (!host_out || n < 0 || n > 16) return EDA_ERR_INVALID_ARGUMENT;
  EDA_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_wg_ts, sizeof(long long) * n), "wgrad debug timestamps");
  return EDA_OK;
}

bool wgrad_tc_eligible(const eda_wgrad_problem *probs, int nprobs, int N, int K) {
  static const bool off = [] { const char *e = getenv("EDA_WGRAD_TC"); return e && e[0] == '0'; }();
  if (off || !encode_tiled_fn() || nprobs < 1 || nprobs > kTcMaxProbs) return false;
  if ((N & 3) || (K & 3) || N < 64 || K < 32) return false;
  long long max_rows = 0;
  for (int i = 0; i < nprobs; ++i) {
    const eda_wgrad_problem &q = probs[i];
    if ((q.x_scale == nullptr) != (q.x_shift == nullptr)) return false;
    if (q.rows <= 0 || q.rows > 0x7fffffffLL - kTcRows) return false;
    if ((q.ldy & 3) || (q.ldx & 3) || (reinterpret_cast<uintptr_t>(q.dy) & 15) || (reinterpret_cast<uintptr_t>(q.x) & 15))
      return false;
    if (q.rows > max_rows) max_rows = q.rows;
  }
  // below a few thousand rows in total the fixed cost of a CTA (set-up, first TMA round trip, 128 x K reduction into dW)
  // outweighs the faster contraction: 640 rows x 1 problem 6.6 us (warp-level) vs 9.2 us, 2048 x 3 22.5 vs 11.2 us.
  // EDA_WGRAD_TC=2 takes every eligible shape (tests), =0 none.
  long long total_rows = 0;
  for (int i = 0; i < nprobs; ++i) total_rows += probs[i].rows;
  static const bool force = [] { const char *e = getenv("EDA_WGRAD_TC"); return e && e[0] == '2'; }();
  return max_rows >= 256 && (force || total_rows >= 3000);
}

int wgrad_tc_launch(const eda_wgrad_problem *probs, int nprobs, int N, int K, cudaStream_t stream) {
  TcParams p = {};
  long long max_rows = 0;
  for (int i = 0; i < nprobs; ++i) {
    const eda_wgrad_problem &q = probs[i];
    if (!make_tensor_map_rows32(&p.map_y[i], q.dy, q.rows, N, q.ldy, kTcRows, true) ||
        !make_tensor_map_rows32(&p.map_x[i], q.x, q.rows, K, q.ldx, kTcRows, true))
      return kWgradTcDeclined;
    p.pr[i].dw = q.dw; p.pr[i].db = q.db; p.pr[i].rows = q.rows; p.pr[i].ldw = q.ldw;
    p.pr[i].x_scale = q.x_scale; p.pr[i].x_shift = q.x_shift;
    p.pr[i].vec = ((q.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(q.dw) & 15) == 0) ? 1 : 0;
    if (q.rows > max_rows) max_rows = q.rows;
  }
  p.nprobs = nprobs; p.N = N; p.K = K;
  p.ntiles = (N + kTileN - 1) / kTileN;
  p.kchunks = (K + kTcMaxChunk - 1) / kTcMaxChunk;
  p.kchunk = (((K + p.kchunks - 1) / p.kchunks) + 31) & ~31;
  const int nb = ((p.kchunk < K ? p.kchunk : K) + 31) / 32;  // X boxes per stage (<= kTcMaxBoxes)
  p.stage_bytes = (uint32_t)((kYBoxes + nb) * kBoxBytes);
  p.nstages = (int)((216 * 1024) / p.stage_bytes);
  if (p.nstages > kTcMaxStages) p.nstages = kTcMaxStages;
  if (p.nstages < 2 || nb > kTcMaxBoxes) return kWgradTcDeclined;
  // one CTA per SM (the ring takes most of the shared memory): as many row slices as fill the chip once, at least two
  // 32-row stages each
  const long long units = (long long)p.ntiles * p.kchunks * nprobs;
  const long long chunks = (max_rows + kTcRows - 1) / kTcRows;
  // EDA_WGRAD_TC_SMS caps the CTAs of one launch (measurements: the kernel runs on a side stream next to the step's
  // critical path, whose kernels cannot share an SM with a 200 KB CTA)
  static const int sm_cap = [] { const char *e = getenv("EDA_WGRAD_TC_SMS"); return e ? atoi(e) : 0; }();
  const int sm_budget = (sm_cap > 0 && sm_cap < sm_count()) ? sm_cap : sm_count();
  long long splits = sm_budget / units;
  if (splits > chunks / 2) splits = chunks / 2;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  // the epilogue reuses the ring: bias scratch + the accumulator tile, padded (all columns when they fit next to the
  // ring's size, else one MMA piece of <= 256 columns at a time)
  size_t smem = (size_t)p.nstages * p.stage_bytes;
  const size_t ep_all = kScratchBytes + (size_t)kTileN * (32 * nb + 4) * sizeof(float);
  const size_t ep_piece = kScratchBytes + (size_t)kTileN * ((nb > 8 ? 256 : 32 * nb) + 4) * sizeof(float);
  p.ep_whole = ep_all <= (smem > ep_piece ? smem : ep_piece) ? 1 : 0;
  if (smem < ep_piece) smem = ep_piece;
  smem += 1024;
  static SmemAttr attr;
  EDA_CUDA_TRY(attr.ensure(wgrad_tc_kernel, smem), "wgrad_tc smem attr");
  EDA_CUDA_TRY(launch_pdl(wgrad_tc_kernel, dim3((unsigned)(units * splits)), dim3(kTcThreads), smem, stream, p),
               "wgrad_tc_kernel launch");
  return check_launch("wgrad_tc_kernel");
}

}  // namespace eda
