// tcgen05 (5th-gen tensor core) / TMEM primitives for sm_100a, written as inline PTX.
//
// Conventions used by every kernel in this directory:
//   * kind::tf32 operands: fp32 words in shared memory (or TMEM for A), fp32 accumulate in TMEM.
//   * Shared-memory operands use the K-major, SWIZZLE_NONE canonical layout in "chunk-major" form:
//         element (row r, k) lives at   base + (k/4) * LBO + (r/8) * 128 + (r%8) * 16 + (k%4) * 4
//     i.e. an array  float4 tile[K/4][rows]  (LBO = rows*16 bytes, SBO = 128 bytes).  A core
//     matrix is 8 rows x 16 bytes, stored contiguously; one MMA consumes K = 8 (two 16-byte
//     chunks, LBO apart).  Row-contiguous float4 stores by 32 consecutive threads are
//     bank-conflict free in this layout.
//   * Accumulator / TMEM-A layout for M = 128, cta_group::1: TMEM lane = row, column = n (or k).
#pragma once
#include <stdint.h>
#include "common.cuh"

namespace eda {
namespace umma {

// ---- descriptors --------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (sm_100 format: version field = 1).
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes,
                                                               uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);         // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;   // [16,30) leading (K) byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;   // [32,46) stride (M/N, per 8 rows) byte offset >> 4
  d |= (uint64_t)1 << 46;                              // [46,48) descriptor version = 1 (Blackwell)
  // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
  return d;
}

// Same descriptor with a swizzle mode: layout_type 0 = none, 1 = 128B (base 32B), 2 = 128B, 4 = 64B, 6 = 32B.
// K-major SWIZZLE_128B: rows are 128 bytes (32 tf32), 8-row atoms of 1024 bytes (1024-byte aligned), the
// 16-byte chunk j of row r sits at chunk position j ^ (r & 7); SBO = byte distance between 8-row atoms,
// LBO unused; advancing K by 8 elements = +32 bytes on the start address.
__device__ __forceinline__ uint64_t smem_desc_swizzled(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                       uint32_t layout_type) {
  return smem_desc_kmajor_noswizzle(smem_addr, lbo_bytes, sbo_bytes) | ((uint64_t)(layout_type & 7u) << 61);
}

// 32-bit instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4)                    // c_format = F32
         | (2u << 7)                  // a_format = TF32
         | (2u << 10)                 // b_format = TF32
         | (0u << 15) | (0u << 16)    // a_major = K, b_major = K
         | ((uint32_t)(N >> 3) << 17) // n_dim
         | ((uint32_t)(M >> 4) << 24);// m_dim
}

// Same, B operand MN-major ("transposed").  Measured on B200 (scripts/probe_umma_mn.py): kind::tf32 reads MN-major
// operands only in the 128-byte swizzle with 32-byte atoms (descriptor layout type 1: 128-byte rows of 32 consecutive
// MN elements per k, 32-byte units XOR (k & 3), 4 k rows per 512-byte group = SBO, next 32 MN elements = LBO) —
// the no-swizzle and 16-byte-atom layouts return zeros.  See csrc/wgrad_tc.cu for the user.
__host__ __device__ constexpr uint32_t idesc_tf32_bmn(int M, int N) { return idesc_tf32(M, N) | (1u << 16); }

// ---- Ampere-style asynchronous 16-byte copies global -> shared (LDGSTS): many in flight, no registers.
// src_bytes < 16 zero-fills the remainder (0 = pure zero fill, nothing is read).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMEM allocation ------------------------------------------------------------------------
// One full warp calls alloc; the base address lands in *smem_slot.  ncols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- MMA issue (one thread) -------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T          A: 128 x 8 (K-major), B: N x 8 (K-major)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T          A: 128 lanes x 8 columns of fp32
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-converged forms: the WHOLE warp executes the call with warp-uniform arguments and one elected lane issues.  An
// issue loop written this way keeps its descriptors in uniform registers (computed by the uniform datapath); the same
// loop inside `if (lane == 0)` makes the compiler move every descriptor from vector to uniform registers through an
// elect / R2UR / branch "waterfall" per MMA, which costs more than the MMA itself (eda_selftest_umma_rate: 161 -> 72
// cycles per 128 x 144 x 8 MMA).
__device__ __forceinline__ void mma_tf32_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four consecutive K steps (K = 8 each) into one accumulator from ONE statement: step k reads descriptors whose low
// words are a_lo + k * a_step / b_lo + k * b_step (the 14-bit "address >> 4" field lives in the low word; the high
// words do not change).  One elect and one branch per group, two adds per MMA: the tensor pipe, not the issue loop,
// sets the pace (72 cycles per 128 x 144 x 8 MMA instead of 130 - 160 with per-MMA issue code).
__device__ __forceinline__ void mma4_tf32_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t a_step, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q, t;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 al, bl;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@!q bra MMA4_SKIP;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 t, %0, %0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %7, p;\n\t"
      "add.u32 al, %1, %3;\n\t"
      "add.u32 bl, %4, %6;\n\t"
      "mov.b64 da, {al, %2};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %7, t;\n\t"
      "add.u32 al, al, %3;\n\t"
      "add.u32 bl, bl, %6;\n\t"
      "mov.b64 da, {al, %2};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %7, t;\n\t"
      "add.u32 al, al, %3;\n\t"
      "add.u32 bl, bl, %6;\n\t"
      "mov.b64 da, {al, %2};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %7, t;\n\t"
      "MMA4_SKIP:\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(a_step), "r"(b_lo), "r"(b_hi), "r"(b_step), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// Same with the A operand in tensor memory (a_tmem advances by a_step columns per K step).
__device__ __forceinline__ void mma4_tf32_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t a_step, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q, t;\n\t"
      ".reg .b64 db;\n\t"
      ".reg .b32 al, bl;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@!q bra MMA4_SKIP;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 t, %0, %0;\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %6, p;\n\t"
      "add.u32 al, %1, %2;\n\t"
      "add.u32 bl, %3, %5;\n\t"
      "mov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], db, %6, t;\n\t"
      "add.u32 al, al, %2;\n\t"
      "add.u32 bl, bl, %5;\n\t"
      "mov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], db, %6, t;\n\t"
      "add.u32 al, al, %2;\n\t"
      "add.u32 bl, bl, %5;\n\t"
      "mov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], db, %6, t;\n\t"
      "MMA4_SKIP:\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(a_step), "r"(b_lo), "r"(b_hi), "r"(b_step), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return (uint32_t)(d >> 32); }
// Warp index as a value the compiler KNOWS is warp-uniform (threadIdx.x >> 5 alone is not): role branches on it keep the
// code inside on the uniform datapath.
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void mma_commit_w(uint64_t *bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// All MMAs issued so far by this thread arrive (once) on `bar` when they complete.  Implies
// tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM <-> registers (warp-wide; warp w touches lanes [32*(w%4), 32*(w%4)+32)) -------------
// 32 lanes x 16 consecutive columns: thread t gets lane base+t, columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address = (lane << 16) | column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

}  // namespace umma
}  // namespace eda
