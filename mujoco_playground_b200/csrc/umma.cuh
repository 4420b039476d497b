// umma.cuh -- thin inline-PTX layer over the sm_100a tensor-core path (tcgen05.mma with TMEM accumulators, mbarrier completion),
// for kind::tf32 operands held in shared memory in the canonical 128-byte-swizzled layouts.
//
// Shared-memory operand layout used throughout ("SW128 block"): a block holds R rows of 128 bytes (32 fp32); rows are grouped in
// 8-row swizzle atoms of 1024 bytes; inside an atom the 16-byte chunk c of row r is stored at chunk position c ^ (r & 7)
// (Swizzle<3,4,3> on the byte address; the block base must be 1024-byte aligned).  Element (r, c) of a block at `base`:
//     base + (r >> 3) * 1024 + (r & 7) * 128 + (((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4
// The SAME bytes serve two descriptor interpretations:
//   * K-major operand   [MN = rows r][K = 32 columns of the block]           (an [M x K] activation tile as A, an [N x K] weight as B)
//   * MN-major operand  [K = rows r][MN = 32 columns of the block, further MN blocks LBO bytes apart]
// so an activation array stored once as [sample][feature] is the A operand of the forward / backward-data GEMMs (K = feature) and
// the A / B operand of the weight-gradient GEMMs (K = sample) without a transposed copy.
#pragma once
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, col) inside one SW128 block (col < 32)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) & 7) << 4) + (c & 3) * 4);
}

// 64-bit shared-memory matrix descriptor (SM100 UMMA): start address, leading / stride byte offsets (all >> 4), version 1,
// layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// K-major operand: 8-row atoms 1024 bytes apart; the K extent of one instruction (8 x tf32 = 32 bytes) lies inside the 128-byte row,
// successive K steps advance the start address by 32 bytes
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t block_saddr, int kstep_in_block) {
  return make_desc(block_saddr + 32u * (uint32_t)kstep_in_block, 16u, 1024u);
}
// MN-major operand in the plain SWIZZLE_128B layout: MN blocks (32 elements each) `block_stride` bytes apart, one K step = 8 rows = one
// 1024-byte atom.  Valid for 16-bit operands only: for tf32 the tensor core returns zeros (tools/microbench/umma_probe.cu shows it);
// the kernels use desc_mn32 / mn32_lo (SWIZZLE_128B_BASE32B) below.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t block0_saddr, uint32_t block_stride, int kstep) {
  return make_desc(block0_saddr + 1024u * (uint32_t)kstep, block_stride, 1024u);
}

// instruction descriptor, kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// same with the A operand in TMEM (128 lanes = rows, one 32-bit column per tf32 element; a_tmem = first column of the K step)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// The same two instructions with the descriptors given as (low word, high word): the low word holds the start address (>> 4) and the
// leading offset, the high word is the same for every K step of an operand -- a K step is then ONE 32-bit add on the low word in the
// issuing thread (the 64-bit shift/mask construction per instruction costs the single issuing thread ~80 cycles per MMA).
__device__ __forceinline__ void mma_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptor words.  K-major SWIZZLE_128B operand (LBO 16 B, SBO 1024 B): K step k of a chain of 8192-byte blocks of 64 rows adds
// (k >> 2) * 512 + (k & 3) * 2 to the low word.  MN-major SWIZZLE_128B_BASE32B operand (LBO = block stride, SBO 512 B): K step k adds 64 k.
__device__ __forceinline__ uint32_t kmajor_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t KMAJOR_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t mn32_lo(uint32_t saddr, uint32_t block_stride) { return ((saddr >> 4) & 0x3FFFu) | (((block_stride >> 4) & 0x3FFFu) << 16); }
constexpr uint32_t MN32_HI = (512u >> 4) | (1u << 14) | (1u << 29);

// one lane of a converged warp (the same lane every time for the full mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// all MMAs issued so far by this thread arrive on the mbarrier when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared, cp.async) -> visible to the async proxy (tensor-core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(mbar)), "r"(parity)
      : "memory");
}

// bulk copy global -> shared memory (async proxy), completion counted in bytes on an mbarrier of this CTA
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src), "r"(bytes),
               "r"(smem_u32(mbar))
               : "memory");
}

// TMEM allocation: one warp allocates `cols` (power of two >= 32) columns and publishes the base address in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quarter: lane base = 32 * (warp % 4) in bits [31:16] of taddr) x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  // the wait carries the registers as in/out operands so that no use of them can be scheduled above it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: this thread's lane, 16 consecutive columns (warp-collective; completes at tmem_st_wait)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
      "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B_BASE32B: the one MN-major layout the tensor core accepts for 32-bit (tf32) operands.  A block holds rows (= K index) of
// 128 bytes (32 MN elements); 4-row atoms of 512 bytes; the 32-byte chunk c of row r is stored at chunk position c ^ (r & 3).
__device__ __forceinline__ uint32_t b32_off(int r, int c) {
  return (uint32_t)((r >> 2) * 512 + (r & 3) * 128 + ((((c >> 3) ^ (r & 3)) & 3) << 5) + (c & 7) * 4);
}
// MN-major tf32 operand: MN blocks `block_stride` bytes apart (LBO), 4-row atoms 512 bytes apart (SBO); one K step = 8 rows = 1024 bytes
__device__ __forceinline__ uint64_t desc_mn32(uint32_t block0_saddr, uint32_t block_stride, int kstep) {
  uint64_t d = 0;
  d |= (uint64_t)(((block0_saddr + 1024u * (uint32_t)kstep) >> 4) & 0x3FFF);
  d |= (uint64_t)((block_stride >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                       // SWIZZLE_128B_BASE32B
  return d;
}

}  // namespace umma
